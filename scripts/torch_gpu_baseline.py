#!/usr/bin/env python
"""Informative second baseline: what the reference's own code path gives on a B200 -- the same two
networks as plain torch.nn modules running through cuDNN / cuBLAS (bf16 autocast + channels_last, and fp32 with TF32),
forward + loss + backward, same batch sizes as bench.py.  Not a product path and not the parity oracle: the modules are
restated here from the reference's layer lists (model.py:88-209, vae-kl.ipynb:122-166) only to time stock PyTorch.

    python scripts/torch_gpu_baseline.py [--steps 20]
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):                     # model.py:23-55
    def __init__(self, cin, c):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(cin, c, 3, 2, 1, bias=False), nn.BatchNorm2d(c)
        self.conv2, self.bn2 = nn.Conv2d(c, c, 3, 1, 1, bias=False), nn.BatchNorm2d(c)
        self.down = nn.Sequential(nn.Conv2d(cin, c, 1, 2, bias=False), nn.BatchNorm2d(c))

    def forward(self, x):
        o = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(o)) + self.down(x))


class UpBlock(nn.Module):                        # model.py:57-85
    def __init__(self, cin, c):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(cin, c, 1, bias=False), nn.BatchNorm2d(c)
        self.conv2, self.bn2 = nn.ConvTranspose2d(c, c, 4, 2, 1, bias=False), nn.BatchNorm2d(c)
        self.up = nn.Sequential(nn.ConvTranspose2d(cin, c, 4, 2, 1, bias=False), nn.BatchNorm2d(c))

    def forward(self, x):
        o = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(o)) + self.up(x))


class ResnetVAE(nn.Module):                      # model.py:88-209, 64x64, z = 64
    def __init__(self, z=64, w=1):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(1, 32 * w, 5, 2, 2, bias=False), nn.BatchNorm2d(32 * w), nn.ReLU())
        self.enc = nn.Sequential(BasicBlock(32 * w, 32 * w), BasicBlock(32 * w, 64 * w), BasicBlock(64 * w, 128 * w),
                                 BasicBlock(128 * w, 256 * w), nn.AdaptiveAvgPool2d(1))
        self.mu, self.lv = nn.Conv2d(256 * w, z, 1, bias=False), nn.Conv2d(256 * w, z, 1, bias=False)
        self.dstem = nn.Sequential(nn.ConvTranspose2d(z, 128 * w, 2, bias=False), nn.BatchNorm2d(128 * w), nn.ReLU())
        self.dec = nn.Sequential(UpBlock(128 * w, 128 * w), UpBlock(128 * w, 64 * w), UpBlock(64 * w, 32 * w),
                                 UpBlock(32 * w, 16 * w), UpBlock(16 * w, 16 * w))
        self.tail = nn.Sequential(nn.Conv2d(16 * w, 1, 3, 1, 1), nn.BatchNorm2d(1))

    def step(self, x, _y):
        h = self.enc(self.stem(x))
        mu, lv = self.mu(h), self.lv(h)
        z = mu + torch.randn_like(mu) * torch.exp(0.5 * lv)
        r = self.tail(self.dec(self.dstem(z))).float()
        n = x.shape[0]
        nll = ((x - r) ** 2 / (2 * 0.01)).sum() + x.numel() * (-2.3025851 + 0.9189385)
        kl = -0.5 * torch.sum(lv.float() - lv.float().exp() - mu.float() ** 2 + 1)
        return (nll + kl) / n


class NotebookVAE(nn.Module):                    # vae-kl.ipynb:122-166
    def __init__(self, c=32, z=32):
        super().__init__()
        self.e = nn.ModuleList([nn.Conv2d(1, c, 5, 2, 2), nn.Conv2d(c, c, 5, 2, 1), nn.Conv2d(c, c, 3, 2, 1), nn.Conv2d(c, c, 3, 2, 1)])
        self.mu, self.lv = nn.Conv2d(c, z, 3, 2, 1), nn.Conv2d(c, z, 3, 2, 1)
        self.d = nn.ModuleList([nn.Conv2d(z, c, 3, 1, 1), nn.Conv2d(c, c, 3, 1, 1), nn.Conv2d(c, c, 3, 1, 1), nn.Conv2d(c, 256, 3, 1, 1)])

    def step(self, x, y):
        h = x
        for conv in self.e:
            h = F.relu(conv(h))
        mu, lv = self.mu(h), self.lv(h)
        h = mu + torch.randn_like(mu) * torch.exp(0.5 * lv)
        for conv, f in zip(self.d[:3], (2, 4, 2)):
            h = F.elu(conv(F.interpolate(h, scale_factor=f, mode="nearest")))
        logits = self.d[3](F.interpolate(h, scale_factor=2, mode="nearest"))
        n = x.shape[0]
        pxz = (F.cross_entropy(logits.float(), y, reduction="none") / n).sum()
        kl = -0.5 * torch.sum(lv.float() - lv.float().exp() - mu.float() ** 2 + 1) / n
        return pxz + kl


def time_model(model, x, y, steps, autocast, graph=False):
    params = list(model.parameters())

    def one():
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = model.step(x, y)
        loss.backward()

    if graph:
        # the same step captured once and replayed (no host launch overhead): the fair comparison with our graph replay
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                one()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            one()
        one = g.replay                                 # noqa: F811
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    for name, mk, n, size in (("model.py VAE, 256 x 64x64", lambda: ResnetVAE(), 256, 64),
                              ("model.py VAE widened, 128 x 64x64", lambda: ResnetVAE(256, 2), 128, 64),
                              ("vae-kl.ipynb VAE, 512 x 128x128", lambda: NotebookVAE(), 512, 128)):
        x = torch.randn(n, 1, size, size, device=dev, generator=g)
        y = torch.randint(0, 256, (n, size, size), device=dev, generator=g)
        for mode, autocast, cl, graph in (("bf16 autocast, channels_last", True, True, False),
                                          ("bf16 autocast, channels_last, CUDA-graph replay", True, True, True),
                                          ("fp32 (TF32 allowed)", False, False, False)):
            torch.manual_seed(0)
            m = mk().to(dev).train()
            xx = x
            if cl:
                m = m.to(memory_format=torch.channels_last)
                xx = x.contiguous(memory_format=torch.channels_last)
            try:
                ms = time_model(m, xx, y, args.steps, autocast, graph)
                out[f"{name}; {mode}"] = {"ms_per_step": ms, "frames_per_s": n / ms * 1e3}
            except RuntimeError as e:                    # e.g. out of memory for the fp32 logits
                out[f"{name}; {mode}"] = {"error": str(e).split("\n")[0][:160]}
            del m
            torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
