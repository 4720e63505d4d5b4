#!/usr/bin/env python
"""Throughput of the notebook-variant VAE step (BASELINE configs[4]: 128x128 frames, 512 per GPU, bf16) on one B200,
and the per-kernel breakdown of one step (CUPTI through torch.profiler -- a diagnostic, not a bench number).

    python scripts/bench_nb.py [--batch 512] [--size 128] [--steps 20] [--profile]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mmvae_b200 as M  # noqa: E402


def synthetic(n, size, dev, seed=1234):
    """grey-level frames 0..255 with two bright blobs on a black background, generated on the device"""
    g = torch.Generator(device=dev).manual_seed(seed)
    y = torch.zeros(n, size, size, dtype=torch.int64, device=dev)
    d = 28 * size // 64
    oy = torch.randint(0, size - d + 1, (n, 2), generator=g, device=dev)
    ox = torch.randint(0, size - d + 1, (n, 2), generator=g, device=dev)
    blob = (torch.rand(n, 2, d, d, generator=g, device=dev) * 255).long()
    for i in range(n):
        for b in range(2):
            y[i, oy[i, b]:oy[i, b] + d, ox[i, b]:ox[i, b] + d] = blob[i, b]
    x = ((y.float() / 255.0 - 0.1307) / 0.3081).unsqueeze(1).contiguous()
    return x, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = M.NotebookVAE(1, 32, 32, image_size=args.size, precision=args.precision).to(dev)
    n = args.batch
    x, y = synthetic(n, args.size, dev)
    info = M._lib.layout(m._desc(n))
    for _ in range(args.warmup):
        out = m.train_step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = M._lib.lib.mmvae_launch_count()
    e0.record()
    for _ in range(args.steps):
        out = m.train_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (M._lib.lib.mmvae_launch_count() - l0) // args.steps
    fps = n / (ms * 1e-3)
    tf = info.train_flops / (ms * 1e-3) / 1e12
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    sustained = peaks.get("bf16_tflops_sustained", 1399.0)
    print(json.dumps({"workload": f"vae-kl.ipynb VAE {args.size}x{args.size}, {n} frames, {args.precision}, fwd+loss+bwd",
                      "ms_per_step": ms, "frames_per_s": fps, "tflops": tf, "frac_of_sustained_bf16_peak": tf / sustained,
                      "launches_per_step": launches, "workspace_gb": info.workspace_bytes / 1e9,
                      "loss": [float(v) for v in out.cpu()]}))
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m.train_step(x, y)
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        print(f"{'start':>9} {'dur us':>9}  kernel")
        for e in evs:
            print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:9.1f}  {e.name[:110]}")
        print(f"span {evs[-1].time_range.end - t0:.1f} us")


if __name__ == "__main__":
    main()
