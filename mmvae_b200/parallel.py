"""Data parallelism for the VAE step: one process per GPU, frames sharded across ranks, one exchange
step -- the gradient all-reduce (the reference itself is single-device, main.py:433-437).

BatchNorm statistics stay per replica (the reference has no SyncBN), so the result is the mean over
ranks of the per-shard gradients.  The backward sweep is issued in three phases (decoder, deep encoder,
shallow encoder: include/mmvae.h MMVAE_BWD_*); each phase owns a contiguous range of the flat gradient
arena, and that range's NCCL all-reduce (ReduceOp.AVG over NVLink/NVSwitch) starts on a side stream as
soon as the phase's last kernel is done, overlapping the rest of the backward.
"""
import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_items, rank, world):
    """[begin, end) of the contiguous shard of `n_items` frames owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradSync:
    """Bucketed, overlapped gradient averaging over a process group."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.comm_stream = None
        self._ranges = {}
        self.bytes_reduced = 0
        # phase_done() makes the communication stream wait for the library's auxiliary streams itself (mmvae_aux_fence),
        # so mmvae_backward need not join a phase's weight gradients into the compute stream (MMVAE_BWD_DEFER_JOIN)
        self.fences_aux = torch.cuda.is_available()

    def ranges(self, desc):
        key = (desc.batch, desc.width, desc.z_dim, desc.image_size, desc.out_channels, desc.in_channels)
        if key not in self._ranges:
            self._ranges[key] = {ph: _lib.backward_range(desc, ph)
                                 for ph in (_lib.BWD_DECODER, _lib.BWD_ENC_DEEP, _lib.BWD_ENC_SHALLOW)}
        return self._ranges[key]

    def reduce_range(self, flat, begin, end):
        """Average flat[begin:end] over the group (in place)."""
        seg = flat[begin:end]
        if self.world == 1:
            return
        if flat.is_cuda:
            dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo has no AVG
            dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
            seg.div_(self.world)
        self.bytes_reduced += seg.numel() * seg.element_size()

    def phase_done(self, model, desc, grads, phase):
        begin, end = self.ranges(desc)[phase]
        if not grads.is_cuda:
            self.reduce_range(grads, begin, end)
            return
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()                             # the phase's last kernel, on the compute stream
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            if self.fences_aux and getattr(self, "fence_now", False):
                import ctypes
                _lib.check(_lib.lib.mmvae_aux_fence(ctypes.c_void_p(self.comm_stream.cuda_stream)), "mmvae_aux_fence")
            self.reduce_range(grads, begin, end)
        if not torch.cuda.is_current_stream_capturing():
            grads.record_stream(self.comm_stream)

    def finish(self):
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


def data_parallel(model, group=None, broadcast=True):
    """Attach gradient averaging to a mmvae_b200.VAE (three overlapped buckets) or a mmvae_b200.NotebookVAE (its
    165 k parameters are one 0.66 MB bucket, all-reduced once after the backward call) and make the replicas start
    from rank 0's weights."""
    sync = GradSync(group)
    if broadcast and sync.world > 1:
        dist.broadcast(model.flat_parameters, src=0, group=group)
        if getattr(model, "_bn_arena", None) is not None and model._bn_arena.numel():
            dist.broadcast(model._bn_arena, src=0, group=group)
    model._grad_sync = sync if sync.world > 1 else None
    # decorrelate the rsample noise across ranks: same seed, disjoint Philox counter ranges
    if dist.is_initialized():
        model._philox_offset += dist.get_rank(group) * (1 << 40)
    return model
