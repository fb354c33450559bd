"""CPU: host-side logic of the drop-in — event queue timing (F8), synth() segmentation,
voice partitioning, and the sharded N>1 path over gloo with the CPU restatement standing in
for the engine (tests only)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import cases
from oracle import oracle as O
from skred_b200 import workloads as W
from tests_util import assert_state_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not O.have_ref(64), reason="compiled reference (oracle/_ref) not present")


def queue(s, wl):
    ev = W.to_skb_events(wl["timed"])
    s.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    assert s.lib.skb_shim_queue_events(ev.ctypes.data, len(ev)) == 0
    return ev


def test_callback_rule():
    # seq() after the callback ending at c fires when <= c + 512 (seq.c:171-178)
    assert W.callback_for_time(0) == 1 and W.callback_for_time(1024) == 1
    assert W.callback_for_time(1025) == 2 and W.callback_for_time(30 * 44100) == 2583


@pytest.mark.parametrize("block", [512, 4096, 1536])
def test_event_queue_equals_harness_injection(block, luts):
    """Timestamped queue inside the shim == events injected between 512-frame callbacks,
    also when the host asks for multi-block calls (synth() is cut where something fires)."""
    wl = cases.SYNTHETIC["pcm_retrigger"](luts)
    a = O.RefSkred(64, run_seq=False) if O.have_ref(64) else O.PortSkred(64, run_seq=False)
    b = O.PortSkred(64, run_seq=False)
    for s in (a, b):
        cases.drive_setup(s, wl)
    oa = a.render(wl["frames"], events=wl["events"])
    queue(b, wl)
    ob = b.render(wl["frames"], block=block)
    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert_state_equal(a.state(), b.state())
    assert b.lib.skb_shim_pending_events() == sum(1 for w, _ in wl["timed"] if W.callback_for_time(w) * 512 >= wl["frames"])


def test_workload_generators_are_deterministic(luts):
    for fn in (lambda: W.config2(64, 2.0, luts), lambda: W.config3(64, 1.0, 4), lambda: W.config4(64, 0.5),
               lambda: W.config5(96, 20.0, luts)):
        x, y = fn(), fn()
        assert x["setup"] == y["setup"] and x["events"] == y["events"]
    wl = W.config5(96, 30.0, luts)
    assert len(wl["timed"]) == 96 * 3                 # one (re)trigger per voice per 10 s


# ---- partitioner ---------------------------------------------------------------
class Params(C.Structure):
    _fields_ = [("amp", C.c_float), ("phase_inc", C.c_float), ("freq_scale", C.c_float), ("freq_mod_depth", C.c_float),
                ("freq_mod_osc", C.c_int32), ("table_id", C.c_int32), ("table_size", C.c_int32),
                ("loop_start_f", C.c_float), ("loop_end_f", C.c_float), ("flags", C.c_uint32), ("cz_mode", C.c_int32),
                ("cz_distortion", C.c_float), ("cz_mod_osc", C.c_int32), ("cz_mod_depth", C.c_float),
                ("sample_hold_max", C.c_int32), ("quantize", C.c_int32), ("filter_mode", C.c_int32),
                ("b0", C.c_float), ("b1", C.c_float), ("b2", C.c_float), ("a1", C.c_float), ("a2", C.c_float),
                ("env_attack", C.c_float), ("env_decay", C.c_float), ("env_sustain", C.c_float), ("env_release", C.c_float),
                ("amp_mod_osc", C.c_int32), ("amp_mod_depth", C.c_float), ("smoother_k", C.c_float),
                ("pan_mod_osc", C.c_int32), ("pan_mod_depth", C.c_float)]


class Cfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("abi", "device", "n_voices", "max_frames", "rank", "world")] + [
        ("flags", C.c_uint32), ("_r", C.c_int32)]


def test_param_record_is_32_words_minus_one():
    assert C.sizeof(Params) == 31 * 4


def owners(n, world, edges):
    """edges: list of (voice, field, modulator)."""
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libskred_port.so"))
    lib.skb_owns_voice.argtypes = [C.c_void_p, C.c_int]
    lib.skb_set_params.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.skb_destroy.argtypes = [C.c_void_p]
    own = np.zeros((world, n), dtype=np.int32)
    for r in range(world):
        e = C.c_void_p()
        assert lib.skb_create(C.byref(e), C.byref(Cfg(1, 0, n, 512, r, world, 0, 0))) == 0
        for v in range(n):
            p = Params(freq_mod_osc=-1, amp_mod_osc=-1, pan_mod_osc=-1, table_id=-1, amp=1.0)
            for (vv, field, m) in edges:
                if vv == v:
                    setattr(p, field, m)
                    if field == "cz_mod_osc":
                        p.cz_mode, p.cz_mod_depth = 1, 0.5
            lib.skb_set_params(e, v, C.byref(p))
        for v in range(n):
            own[r, v] = lib.skb_owns_voice(e, v)
        lib.skb_destroy(e)
    return own


def test_partition_keeps_groups_together_and_balances():
    edges = [(0, "freq_mod_osc", 1), (5, "amp_mod_osc", 9), (9, "pan_mod_osc", 20), (30, "cz_mod_osc", 31),
             (40, "freq_mod_osc", 40)]            # the last is a self reference: no edge
    own = owners(64, 4, edges)
    assert np.all(own.sum(axis=0) == 1)           # every voice has exactly one owner
    who = own.argmax(axis=0)
    assert who[0] == who[1] and who[5] == who[9] == who[20] and who[30] == who[31]
    counts = np.bincount(who, minlength=4)
    assert counts.max() - counts.min() <= 3
    assert np.array_equal(own, owners(64, 4, edges))      # deterministic: ranks agree without talking


def test_default_cz_edge_is_not_a_dependency():
    # voice_reset never touches cz_mod_osc: default osc 0 / depth 0
    lib_edges = []
    own = owners(32, 2, lib_edges)
    assert abs(int(own[0].sum()) - int(own[1].sum())) <= 1


# ---- N > 1 over gloo ---------------------------------------------------------------
def no_regroup(events):
    """Moving a voice to another shard mid-stream needs a state migration (skb_snapshot /
    skb_restore); the sharded tests keep the modulation graph fixed."""
    return {k: [c for c in v if not c[0].endswith("_mod_set")] for k, v in events.items()}


def _gloo_worker(rank, world, port, frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import cases as cs
    from test_host_logic import no_regroup
    from oracle import oracle as OO
    from skred_b200.sharded import ShardedRenderer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    luts = cs.load_luts()
    wl = cs.SYNTHETIC["mods"](luts)
    wl["events"] = no_regroup(wl["events"])
    s = OO.HarnessSkred.__new__(OO.HarnessSkred)
    # configure the shard BEFORE the first setter creates the engine
    lib = C.CDLL(OO.private_copy(OO.port_lib_path(64)))
    lib.skb_shim_configure.argtypes = [C.c_int] * 4
    assert lib.skb_shim_configure(0, rank, world, 8192) == 0
    OO.SynthAPI.__init__(s, lib, 64)
    lib.ref_init()
    lib.skb_shim_render_mix.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    lib.skb_shim_finish.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.skb_shim_engine.restype = C.c_void_p

    class Api:
        pass
    api = Api()
    api.lib = lib
    api.render_mix = lambda n, ptr, st: lib.skb_shim_render_mix(n, ptr, None)
    api.finish = lambda ptr, n, out, st: lib.skb_shim_finish(ptr, n, out.ctypes.data, 2, None)
    cs.drive_setup(s, wl)
    r = ShardedRenderer(api, dist, device="cpu")
    outs = []
    for k in range(frames // 512):
        if k in wl["events"]:
            s.apply(wl["events"][k])
        o = r.render(512)
        if rank == 0:
            outs.append(o.copy())
    port = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libskred_port.so"))
    port.skb_owns_voice.argtypes = [C.c_void_p, C.c_int]
    owned = [v for v in range(64) if port.skb_owns_voice(C.c_void_p(lib.skb_shim_engine()), v)]
    q.put((rank, np.concatenate(outs) if rank == 0 else None, owned))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_world2_gloo(luts):
    import torch.multiprocessing as mp
    wl = cases.SYNTHETIC["mods"](luts)
    wl["events"] = no_regroup(wl["events"])
    frames = 8 * 512
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    out = res[0][1]
    owned0, owned1 = set(res[0][2]), set(res[1][2])
    assert owned0 | owned1 == set(range(64)) and not (owned0 & owned1)
    for grp in ([0, 1], [5, 7], [10, 11, 12], [20, 21], [52, 53]):          # modulation groups stay on one rank
        assert len({v in owned0 for v in grp}) == 1, grp
    one = O.PortSkred(64, run_seq=False)
    cases.drive_setup(one, wl)
    ref = one.render(frames, events=wl["events"])
    assert np.max(np.abs(out.astype(np.float64) - ref.astype(np.float64))) <= 1e-5


# ---- bench.py: the JSON line at N > 1 (pure dict work) -----------------------------------
def _bench():
    import importlib
    return importlib.import_module("bench")


def test_bench_roofline_objects_are_per_launch():
    b = _bench()
    V, F = 65536, 8192
    k_act = 0.725 * V * F                                   # rendered voice-frames of one launch
    r = b.roofline_objects(k_act, 0.69, V, F, 6446.3, 1965.0, "measured",
                           {"dram_bytes_read": 18541056, "dram_bytes_write": 125184, "warp_instructions": 399726229, "src_hash": "x"})
    alive = k_act / F
    algo = alive * b.BYTES_PER_VOICE_LAUNCH + (V - alive) * 32.0 + F * 8
    # `roofline` is the BINDING bound (fp32 issue); the HBM figure rides beside it
    assert r["roofline"]["bound"] == "fp32-issue" and r["roofline"]["unit"] == "TFLOP/s"
    assert abs(r["roofline"]["frac"] - r["roofline"]["achieved"] / r["roofline"]["peak"]) < 1e-12
    assert abs(r["roofline"]["peak"] - 148 * 128 * 1.965e9 / 1e12) < 1e-9
    assert 0.2 < r["roofline"]["frac"] < 0.6 and 15.0 <= r["roofline"]["flops_per_voice_sample"] <= 32.0
    assert r["roofline"]["traffic"] == 18541056 + 125184 and r["roofline"]["kernel_ms"] == 0.69
    assert r["roofline_fp32"] is r["roofline"]
    assert abs(r["roofline_hbm"]["achieved"] - algo / 0.69e-3 / 1e9) < 1e-9
    assert r["roofline_hbm"]["bound"] == "hbm" and abs(r["roofline_hbm"]["frac"] - r["roofline_hbm"]["achieved"] / 6446.3) < 1e-12
    assert 0.3 < r["roofline_issue"]["frac"] < 0.8
    none = b.roofline_objects(k_act, 0.69, V, F, 6446.3, 1965.0, "measured", None)
    assert none["roofline_issue"] is None and none["roofline"]["traffic"] is None


def test_bench_refuses_stale_ncu_counters(tmp_path):
    """Counters captured from other kernel sources are not reported."""
    import json
    b = _bench()
    p = tmp_path / "c.json"
    p.write_text(json.dumps({"src_hash": "0000", "dram_bytes_read": 1, "dram_bytes_write": 1, "warp_instructions": 1}))
    assert b.load_ncu_counters(str(p)) is None
    p.write_text(json.dumps({"src_hash": b.kernel_source_hash(), "dram_bytes_read": 1, "dram_bytes_write": 1, "warp_instructions": 1}))
    assert b.load_ncu_counters(str(p))["warp_instructions"] == 1


def test_bench_weak_job_is_a_companion_not_the_headline():
    b = _bench()
    V, world, F, steps = 65536, 8, 8192, 5
    roof = b.roofline_objects(0.7 * V // world * F, 0.45, V // world, F, 6446.3, 1965.0, "measured", None)
    line = {"metric": b.METRIC, "value": 7.5e11, "unit": b.UNIT, "n_gpus": world, "steps": steps, "warmup": 3,
            "ms_per_step": 0.52, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": b.WORKLOAD % V, "voices": V, "frames_per_step": F},
            "e2e": {"value": 6.1e11, "unit": b.UNIT, "h2d_bytes_per_step": 1.0, "d2h_bytes_per_step": 2.0},
            "gpu_launches": 15, "clocks": None}
    line.update(roof)
    assert b.weak_companion(line, None, V, world, F, steps) is line
    act = 0.725 * V * world * F * steps
    w_roof = b.roofline_objects(0.725 * V * F, 0.69, V, F, 6446.3, 1965.0, "measured", None)
    weak = {"act": act, "ms": 0.715 * steps, "steps": steps, "launches": 20, "roof": w_roof,
            "e2e": {"value": 3.4e12, "unit": b.UNIT, "h2d_bytes_per_step": 70000.0, "d2h_bytes_per_step": 65536.0, "ms_per_step": 0.9}}
    out = b.weak_companion(line, weak, V, world, F, steps)
    assert out is not line and "weak_scaling" not in line                       # the input is not edited
    # the headline is still BASELINE configs[4] as written: ONE V-voice job
    for k in ("value", "ms_per_step", "scaling", "e2e", "gpu_launches", "config", "roofline"):
        assert out[k] is line[k] or out[k] == line[k], k
    assert out["scaling"] == "strong" and out["config"]["voices"] == V
    w = out["weak_scaling"]
    assert abs(w["value"] - act / (0.715 * steps * 1e-3)) < 1.0 and abs(w["ms_per_step"] - 0.715) < 1e-12
    assert w["voices_total"] == V * world and w["voices_per_gpu"] == V and w["e2e"] == 3.4e12
    assert abs(w["active_fraction"] - 0.725) < 1e-9 and w["roofline"] is w_roof["roofline"]
    import json
    assert "\n" not in json.dumps(out)


def test_bench_l2_flushed_headline():
    b = _bench()
    V, F, steps = 65536, 8192, 20
    line = {"value": 5.5e11, "ms_per_step": 0.70, "scaling": "weak",
            "config": {"voices": V, "l2": "x", "value_l2_flushed": 5.4e11, "active_fraction": 0.72,
                       "value_counting_all_voice_slots": 7.6e11}}
    assert b.l2_flushed_headline(line, 0.0, None, steps, V, F) is line          # no flushed loop (N > 1): untouched
    act, ms = 0.725 * V * F * steps, 0.71 * steps
    out = b.l2_flushed_headline(line, act, ms, steps, V, F)
    assert out is not line and line["value"] == 5.5e11
    assert abs(out["value"] - act / (ms * 1e-3)) < 1.0 and abs(out["ms_per_step"] - 0.71) < 1e-12
    assert out["config"]["value_l2_unflushed"] == 5.5e11 and out["config"]["ms_per_step_l2_unflushed"] == 0.70
    assert "value_l2_flushed" not in out["config"] and out["config"]["l2"].startswith("flushed before every timed step")
    assert abs(out["config"]["active_fraction"] - 0.725) < 1e-12


def test_partition_stable_keeps_owners_and_moves_the_smaller_side(tmp_path):
    """csrc/partition.h: skb_partition_stable.  Ownership survives re-plans: an unchanged component stays, a
    component that merged voices of several shards goes to the shard that held most of them (ties -> lowest rank), a
    component that split leaves its voices where they were.  Compiled here from the header (plain C, shared with the engine)."""
    import subprocess
    src = tmp_path / "pstable.c"
    src.write_text('#include "partition.h"\n'
                   'void t_first(const int32_t *comp, int n, int world, int32_t *owner) { skb_partition(comp, n, world, owner); }\n'
                   'void t_stable(const int32_t *comp, int n, int world, const int32_t *prev, int32_t *owner) '
                   '{ skb_partition_stable(comp, n, world, prev, owner); }\n')
    so = tmp_path / "pstable.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "skred_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                    str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    n, world = 12, 3
    I32 = C.c_int32 * n

    def first(comp):
        o = I32()
        lib.t_first(I32(*comp), n, world, o)
        return list(o)

    def stable(comp, prev):
        o = I32()
        lib.t_stable(I32(*comp), n, world, I32(*prev), o)
        return list(o)

    solo = list(range(n))
    own0 = first(solo)
    assert own0 == [v % 3 for v in range(n)]                       # least-loaded dealing, ties to the lowest rank
    assert stable(solo, own0) == own0                              # nothing changed: nobody moves
    merged = [0, 0, 0, 0] + list(range(4, n))                      # voices 0..3 become one component: owners were 0,1,2,0
    own1 = stable(merged, own0)
    assert own1[:4] == [0, 0, 0, 0] and own1[4:] == own0[4:]       # rank 0 held two of them: the other two move, nobody else does
    tie = [0, 0] + list(range(2, n))                               # voices 0,1 (owners 0 and 1): tie -> lowest rank
    assert stable(tie, own0)[:2] == [0, 0]
    own2 = stable(solo, own1)                                      # the component splits again: its voices stay where they are
    assert own2 == own1
    assert first(merged)[:4] == [0, 0, 0, 0]
