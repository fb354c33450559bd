"""Voice-sharded rendering over the GPUs of one box.

One process per GPU.  Every rank holds the full host mirror of the voice
parameters (setters are cheap and deterministic), its engine owns the voices
the shared partitioner assigns to `rank` — whole modulation groups, never split
— and renders them into a partial stereo mix [frames][2] in device memory.  The
one exchange step of the path is the sum of those partial mixes:
`torch.distributed.reduce(SUM)` to rank 0 (NCCL over NVLink on GPUs; gloo in the
CPU tests), after which rank 0 applies the master volume (synth.c:616-624) and
copies the block to the host.  The summation order across ranks is NCCL's;
within a rank it is fixed.
"""
import ctypes as C

import numpy as np


class ShardedRenderer:
    def __init__(self, api, dist=None, device="cuda", block=512):
        """api: an object with render_mix(n, ptr, stream) / finish(ptr, n, out, stream) and
        .lib.skb_shim_discard_gain (skred_b200.Skred, or the port drop-in in tests)."""
        import torch
        self.torch = torch
        self.api = api
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.device = device
        self.block = block
        self._mix = None
        self.stream = None
        self.tstream = None
        if device != "cpu":
            # a real stream of our own: the engine's kernels and the NCCL reduce are ordered on it
            # (the legacy default stream has handle 0, which the engine reads as "use your own stream")
            self.tstream = torch.cuda.Stream()
            self.stream = C.c_void_p(self.tstream.cuda_stream)

    def _buffer(self, frames):
        if self._mix is None or self._mix.shape[0] < frames:
            self._mix = self.torch.zeros((frames, 2), dtype=self.torch.float32, device=self.device)
        return self._mix[:frames]

    def render_device(self, frames, launch_frames=None):
        """Partial mixes -> summed raw mix on rank 0 (device resident, no host sync)."""
        mix = self._buffer(frames)
        lf = launch_frames or self.block
        ptr = mix.data_ptr()
        for b in range(0, frames, lf):
            n = min(lf, frames - b)
            self.api.render_mix(n, ptr + b * 8, self.stream)
        flush = getattr(self.api.lib, "skb_shim_flush_render", None)
        if flush is not None:
            flush()                      # the engine defers and batches its launches
        if self.dist:
            if self.tstream is not None:
                with self.torch.cuda.stream(self.tstream):
                    self.dist.reduce(mix, dst=0, op=self.dist.ReduceOp.SUM)
            else:
                self.dist.reduce(mix, dst=0, op=self.dist.ReduceOp.SUM)
        return mix

    def render(self, frames, out=None, launch_frames=None):
        """Full path: returns the finished block on rank 0 (None elsewhere)."""
        mix = self.render_device(frames, launch_frames)
        if self.rank == 0:
            if out is None:
                out = np.zeros((frames, 2), dtype=np.float32)
            self.api.finish(mix.data_ptr(), frames, out, self.stream)
            return out
        self.api.lib.skb_shim_discard_gain()
        return None
