"""Synthetic Moving-MNIST-like input for benchmarks and smoke tests (no dataset is available offline).

Mirrors what the reference's loader delivers to the train step: sequences of 20 frames
with two bouncing binary "digit" blobs, flattened to independent frames (movingmnistdataset.py:20-24),
quantised to k=2 k-means labels {0,1} (main.py:25) -- about 5 % ones -- to be normalised by the label
mean / std on the device (main.py:383-388).
"""
import ctypes

import torch

from ._lib import lib, check

DATA_MEAN = 0.0521      # k=2 label statistics, test-output-models.ipynb:40-43
DATA_STD = 0.2222


def synthetic_labels(n_frames, size=64, seq_len=20, seed=1234, device="cpu"):
    """uint8 [n_frames, size, size] label maps in {0,1}; vectorised over sequences."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = max(4, (28 * size) // 64)
    n_seq = (n_frames + seq_len - 1) // seq_len
    frames = torch.zeros(n_seq, seq_len, size, size, dtype=torch.bool)
    yy, xx = torch.meshgrid(torch.arange(d), torch.arange(d), indexing="ij")
    ring = ((yy - (d - 1) / 2.0) ** 2 + (xx - (d - 1) / 2.0) ** 2).sqrt()
    t = torch.arange(seq_len, dtype=torch.float32)
    for _ in range(2):
        r = d * (0.25 + 0.15 * torch.rand(n_seq, generator=g))
        blob = (ring[None] < r[:, None, None]) & (torch.rand(n_seq, d, d, generator=g) < 0.45)
        pos0 = torch.rand(n_seq, 2, generator=g) * (size - d)
        vel = (torch.rand(n_seq, 2, generator=g) - 0.5) * 8.0
        # reflect a straight line into [0, size-d]: triangle wave
        span = float(size - d)
        p = pos0[:, None, :] + vel[:, None, :] * t[None, :, None]
        p = torch.remainder(p, 2 * span)
        p = torch.where(p > span, 2 * span - p, p).long().clamp_(0, size - d)
        for s in range(n_seq):
            for k in range(seq_len):
                py, px = int(p[s, k, 0]), int(p[s, k, 1])
                frames[s, k, py:py + d, px:px + d] |= blob[s]
    out = frames.view(-1, size, size)[:n_frames].to(torch.uint8)
    return out.to(device)


def prepare_input(labels, data_mean=DATA_MEAN, data_std=DATA_STD, want_target=False, out=None, out_target=None):
    """Device side of main.py:381-388: uint8 label map [N,H,W] (CUDA) -> x = (label - mean)/std as
    [N,1,H,W] fp32, and the int64 cross-entropy target when `want_target`."""
    if not labels.is_cuda or labels.dtype != torch.uint8:
        raise ValueError("labels must be a CUDA uint8 tensor")
    labels = labels.contiguous()
    n, h, w = labels.shape
    x = out if out is not None else torch.empty(n, 1, h, w, dtype=torch.float32, device=labels.device)
    if x.shape != (n, 1, h, w) or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("out must be a contiguous fp32 [N,1,H,W] tensor")
    tgt = None
    if want_target:
        tgt = out_target if out_target is not None else torch.empty(n, h, w, dtype=torch.int64, device=labels.device)
    check(lib.mmvae_prepare_input(ctypes.c_void_p(labels.data_ptr()), labels.numel(), data_mean, data_std,
                                  ctypes.c_void_p(x.data_ptr()),
                                  ctypes.c_void_p(tgt.data_ptr() if tgt is not None else 0),
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "mmvae_prepare_input")
    return (x, tgt) if want_target else x


class Prefetcher:
    """Double-buffered host-to-device copies on a copy stream (the input side of the loop, main.py:374-380, as a
    pipeline): `push(host)` starts the copy of the NEXT batch from pinned host memory while the current step runs;
    `pop()` makes the current stream wait for the staged copy and returns the device tensor.  Two staging buffers:
    a buffer is refilled only after the step that read it has been enqueued past its last use (`pop` records that
    point for the buffer handed out by the PREVIOUS pop, i.e. a popped tensor is valid until the next pop)."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]
        self._next = 0
        self._staged = None
        self._out = None

    def push(self, host):
        s = self._next
        self._next ^= 1
        if self.slots[s] is None or self.slots[s].shape != host.shape or self.slots[s].dtype != host.dtype:
            self.slots[s] = torch.empty(host.shape, dtype=host.dtype, device=self.stream.device)
        if self.free[s] is not None:
            self.stream.wait_event(self.free[s])
        with torch.cuda.stream(self.stream):
            self.slots[s].copy_(host, non_blocking=True)
            self.ready[s].record(self.stream)
        self._staged = s

    def pop(self):
        if self._staged is None:
            raise RuntimeError("Prefetcher.pop() without a staged batch: call push(host_tensor) first")
        cur = torch.cuda.current_stream(self.stream.device)
        if self._out is not None:                        # everything enqueued so far has finished with the previous buffer
            ev = torch.cuda.Event()
            ev.record(cur)
            self.free[self._out] = ev
        s, self._staged = self._staged, None
        cur.wait_event(self.ready[s])
        self._out = s
        return self.slots[s]
