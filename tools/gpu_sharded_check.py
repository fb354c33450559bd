#!/usr/bin/env python3
"""N-GPU parity: voice-sharded render (one process per GPU, NCCL reduce of the stereo partial
mixes) against the compiled reference on the host.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
      tools/gpu_sharded_check.py [voices] [blocks]"""
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

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=NB * 512 / 44100.0 + 1.0, stationary=True)
# two modulated pairs so that groups (lock-step bins) are sharded too
wl["setup"] += [("freq_mod_set", 0, 3, 2.0), ("amp_mod_set", 6, 9, 0.5)]
sk = Skred(V, device=local, rank=rank, world=world, max_frames=512)
W.install(sk, wl)
r = ShardedRenderer(sk, dist, device="cuda")
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
if rank == 0:
    from oracle import oracle as O
    ref = O.RefSkred(V, run_seq=False) if O.have_ref(V) else O.PortSkred(V, run_seq=False)
    W.install(ref, wl)
    want = ref.render(NB * 512, events=wl["events"])
    got = np.concatenate(outs)
    d = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    per = [float(np.max(np.abs(got[k * 512:(k + 1) * 512].astype(np.float64) - want[k * 512:(k + 1) * 512]))) for k in range(NB)]
    print("per-block max|diff|:", " ".join("%.2g" % x for x in per[:12]))
    print("got peak per block:", " ".join("%.3g" % float(np.abs(got[k * 512:(k + 1) * 512]).max()) for k in range(6)),
          " want:", " ".join("%.3g" % float(np.abs(want[k * 512:(k + 1) * 512]).max()) for k in range(6)))
    print("sharded x%d: %d voices (%d owned in total), %d frames, max|diff| vs reference %.3g, peak %.3g -> %s" %
          (world, V, int(owned.item()), NB * 512, d, float(np.abs(want).max()), "OK" if d <= 1e-5 and int(owned.item()) == V else "FAIL"))
dist.barrier()
dist.destroy_process_group()
