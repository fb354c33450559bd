#!/usr/bin/env python3
"""N-GPU parity: voice-sharded render (one process per GPU, the exchange step behind the C-ABI: skb_comm_init_rank +
skb_reduce_mix) against the compiled reference on the host, in both cross-rank orders (ncclReduce / rank-ordered
gather + k_sum_ranks), each rendered twice from scratch and compared bit for bit (run-to-run determinism).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
      tools/gpu_sharded_check.py [voices] [blocks] [out.txt]
Exit code 0 = every check passed on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from skred_b200 import Skred, workloads as W  # noqa: E402
from skred_b200.sharded import ShardedRenderer, COMM_NCCL_REDUCE, COMM_ORDERED  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 24
OUT = sys.argv[3] if len(sys.argv) > 3 else None
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=NB * 512 / 44100.0 + 1.0, stationary=True)
# two modulated pairs so that groups (lock-step bins) are sharded too
wl["setup"] += [("freq_mod_set", 0, 3, 2.0), ("amp_mod_set", 6, 9, 0.5)]


def render_once(mode):
    sk = Skred(V, device=local, rank=rank, world=world, max_frames=512, private=True)
    W.install(sk, wl)
    r = ShardedRenderer(sk, dist, device="cuda", comm_mode=mode)
    outs = []
    for k in range(NB):
        if k in wl["events"]:
            sk.apply(wl["events"][k])
        o = r.render(512)
        if rank == 0:
            outs.append(o.copy())
    st = sk.stats()
    owned = torch.tensor([st.n_owned_voices], device="cuda")
    dist.all_reduce(owned)
    dist.barrier()
    if r.eng is not None:
        r.eng.skb_comm_destroy(sk.engine)
    sk.lib.synth_free()
    return (np.concatenate(outs) if rank == 0 else None), int(owned.item())


lines, ok = [], True
res = {}
for name, mode in (("ncclReduce", COMM_NCCL_REDUCE), ("rank-ordered", COMM_ORDERED)):
    a, owned = render_once(mode)
    b, _ = render_once(mode)
    res[name] = (a, b, owned)
if rank == 0:
    from oracle import oracle as O
    ref = O.RefSkred(V, run_seq=False) if O.have_ref(V) else O.PortSkred(V, run_seq=False)
    W.install(ref, wl)
    want = ref.render(NB * 512, events=wl["events"])
    for name, (a, b, owned) in res.items():
        d = float(np.max(np.abs(a.astype(np.float64) - want.astype(np.float64))))
        same = bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
        good = d <= 1e-5 and owned == V and same
        ok &= good
        lines.append("sharded x%d [%s]: %d voices (%d owned in total), %d frames, max|diff| vs %s %.3g, peak %.3g, "
                     "two runs bit-identical: %s -> %s" % (world, name, V, owned, NB * 512, ref.backend, d,
                                                            float(np.abs(want).max()), same, "OK" if good else "FAIL"))
    x = float(np.max(np.abs(res["ncclReduce"][0].astype(np.float64) - res["rank-ordered"][0])))
    lines.append("ncclReduce vs rank-ordered sum: max|diff| %.3g (the order of the cross-rank sum is the only difference)" % x)
    print("\n".join(lines), flush=True)
    if OUT:
        os.makedirs(os.path.dirname(os.path.abspath(OUT)), exist_ok=True)
        with open(OUT, "w") as f:
            f.write("\n".join(lines) + "\n")
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
