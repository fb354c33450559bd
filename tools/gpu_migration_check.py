#!/usr/bin/env python3
"""N-GPU: modulation routing changed AFTER rendering has started (a re-plan that moved a voice to another
shard used to latch an error).  Ownership is sticky now and the state of the smaller side of a merged component
moves to the other GPU inside the re-plan (ncclBroadcast of the records).  Renders a few blocks, joins voices that
live on different ranks with FM / AM / pan-mod edges, renders on, removes one edge again, and compares everything
with the compiled reference on the host.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
      tools/gpu_migration_check.py [out.txt]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from skred_b200 import Skred, workloads as W  # noqa: E402
from skred_b200.sharded import ShardedRenderer  # noqa: E402

OUT = sys.argv[1] if len(sys.argv) > 1 else None
V, NB = 1024, 30
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=NB * 512 / 44100.0 + 1.0, stationary=True)
ev = dict(wl["events"])
# voices 0, 1, 2 ... are dealt round-robin-ish to the ranks by the first plan: these edges join different shards
edits = {6: [("freq_mod_set", 0, 1, 1.5), ("amp_mod_set", 9, 4, 0.4)],
         12: [("pan_mod_set", 21, 2, 0.3), ("freq_mod_set", 3, 0, 0.7)],
         20: [("freq_mod_set", 0, -1, 0.0)]}
for k, calls in edits.items():
    ev[k] = list(ev.get(k, [])) + calls

sk = Skred(V, device=local, rank=rank, world=world, max_frames=512)
W.install(sk, wl)
r = ShardedRenderer(sk, dist, device="cuda")
outs = []
for k in range(NB):
    if k in ev:
        sk.apply(ev[k])
    o = r.render(512)
    if rank == 0:
        outs.append(o.copy())
st = sk.stats()
tot = torch.tensor([st.n_owned_voices, st.migrated_voices], device="cuda", dtype=torch.int64)
dist.all_reduce(tot)
ok = True
if rank == 0:
    from oracle import oracle as O
    ref = O.RefSkred(V, run_seq=False) if O.have_ref(V) else O.PortSkred(V, run_seq=False)
    W.install(ref, wl)
    want = ref.render(NB * 512, events=ev)
    got = np.concatenate(outs)
    d = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    ok = d <= 1e-5 and int(tot[0].item()) == V and int(st.migrated_voices) > 0
    line = ("re-plan with migration x%d: %d voices (%d owned in total), %d voices' state moved between GPUs (per rank %d), "
            "%d frames, max|diff| vs %s %.3g, peak %.3g -> %s" % (world, V, int(tot[0].item()), int(tot[1].item()) // world,
                                                                   int(st.migrated_voices), NB * 512, ref.backend, d,
                                                                   float(np.abs(want).max()), "OK" if ok else "FAIL"))
    print(line, flush=True)
    if OUT:
        with open(OUT, "w") as f:
            f.write(line + "\n")
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.barrier()
if r.eng is not None:
    r.eng.skb_comm_destroy(sk.engine)
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
