"""Voice-sharded rendering over the GPUs of one box.

One process per GPU.  Every rank holds the full host mirror of the voice
parameters (setters are cheap and deterministic), its engine owns the voices
the shared partitioner assigns to `rank` — whole modulation groups, never split
— and renders them into a partial stereo mix [frames][2] in device memory.  The
one exchange step of the path is the sum of those partial mixes, and it lives
BEHIND THE C-ABI: `skb_comm_init_rank` (ncclCommInitRank on the engine's device)
and `skb_reduce_mix` (ncclReduce to rank 0 on the engine's stream, or the
rank-ordered gather + k_sum_ranks: include/skred_b200.h) — a C host reaches N
GPUs without Python.  This module is a caller of that ABI; `torch.distributed`
only ships the 128-byte NCCL id from rank 0 to the other ranks (any transport
would do) and provides the device buffer.  After the sum rank 0 applies the
master volume (synth.c:616-624) and copies the block to the host.

CPU tests (gloo, world_size 2) run the same class over the CPU restatement of
the engine, which has no devices: there the sum is `dist.reduce` — a stand-in
for the exchange step in TEST code paths only (`device="cpu"`).
"""
import ctypes as C
import os

import numpy as np

COMM_ID_BYTES = 128
COMM_NCCL_REDUCE, COMM_ORDERED = 0, 1


def comm_init(api, dist, engine_lib=None, mode=None):
    """Create the engine's NCCL communicator over the ranks of `dist` (one process per GPU):
    rank 0 draws the id (skb_comm_unique_id), the process group ships it, every rank joins
    (skb_comm_init_rank).  Returns the engine library handle."""
    from .host import load_engine_lib
    eng = engine_lib or load_engine_lib()
    eng.skb_comm_unique_id.argtypes = [C.c_void_p]
    eng.skb_comm_init_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    eng.skb_comm_set_mode.argtypes = [C.c_void_p, C.c_int]
    eng.skb_comm_size.argtypes = [C.c_void_p]
    eng.skb_reduce_mix.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    eng.skb_comm_destroy.argtypes = [C.c_void_p]
    rank, world = dist.get_rank(), dist.get_world_size()
    if eng.skb_comm_size(api.engine) == world:
        return eng
    box = [None]
    if rank == 0:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        r = eng.skb_comm_unique_id(buf)
        if r != 0:
            raise RuntimeError("skb_comm_unique_id failed: %d (libnccl.so.2 not found?)" % r)
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0)
    r = eng.skb_comm_init_rank(api.engine, box[0], rank, world)
    if r != 0:
        raise RuntimeError("skb_comm_init_rank failed %d: %s" % (r, eng.skb_error_string(api.engine).decode()))
    if mode is None:
        mode = COMM_ORDERED if os.environ.get("SKB_COMM_ORDERED", "0") == "1" else COMM_NCCL_REDUCE
    eng.skb_comm_set_mode(api.engine, mode)
    return eng


class ShardedRenderer:
    def __init__(self, api, dist=None, device="cuda", block=512, comm_mode=None):
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
        self.eng = None
        if device != "cpu":
            # a real stream of our own: the engine's kernels and the NCCL reduce are ordered on it
            # (the legacy default stream has handle 0, which the engine reads as "use your own stream")
            self.tstream = torch.cuda.Stream()
            self.stream = C.c_void_p(self.tstream.cuda_stream)
            if self.dist:
                self.eng = comm_init(api, self.dist, mode=comm_mode)

    def _buffer(self, frames):
        if self._mix is None or self._mix.shape[0] < frames:
            if self.tstream is not None:
                # allocated and zero-filled ON the stream the kernels and the reduce use: no cross-stream race
                with self.torch.cuda.stream(self.tstream):
                    self._mix = self.torch.zeros((frames, 2), dtype=self.torch.float32, device=self.device)
            else:
                self._mix = self.torch.zeros((frames, 2), dtype=self.torch.float32, device=self.device)
        return self._mix[:frames]

    def reduce(self, mix, frames):
        """The exchange step: partial mixes of all ranks -> their sum in rank 0's `mix` (on self.stream)."""
        if not self.dist:
            return
        if self.eng is not None:
            r = self.eng.skb_reduce_mix(self.api.engine, mix.data_ptr(), frames, self.stream)
            if r != 0:
                raise RuntimeError("skb_reduce_mix failed %d: %s" % (r, self.eng.skb_error_string(self.api.engine).decode()))
        else:
            self.dist.reduce(mix, dst=0, op=self.dist.ReduceOp.SUM)       # CPU restatement under gloo (tests)

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
        self.reduce(mix, frames)
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
