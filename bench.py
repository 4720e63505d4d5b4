#!/usr/bin/env python
"""bench.py -- train frames/sec (forward + loss + backward) of the VAE step on N B200s.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU train step on this box's host cores

Workload (BASELINE.json configs[1]): the reference model.py VAE (z=64, 64x64, Gaussian NLL sigma=0.1),
bf16, 256 frames per GPU of synthetic Moving-MNIST label maps (20-frame sequences flattened to frames),
random-init weights.  N > 1 is weak scaling: 256 frames per GPU, gradients averaged over NVLink.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 256
TRAIN_MFLOP_PER_FRAME = 227.016704      # 2*MAC over every conv, fwd + dgrad + wgrad (mmvae_layout.train_flops)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1399.0), d.get("bf16_tflops", 1608.4), d.get("hbm_gbs", 6527.8), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


REF_DIR = os.path.join(ROOT, "oracle", "_ref")        # the reference's own model.py / vae-kl.ipynb, copied there by
                                                      # __graft_entry__.build() (git-ignored, travels with gpurun)


def _reference_step(batch):
    """The reference's OWN module (oracle/_ref/model.py, unmodified) driven by the loop body main.py:389-390,397-398.
    Returns (step, kind): step() runs one forward + loss + zero_grad + backward; kind 'reference', or 'port' when the
    copy is absent and the oracle port (same torch CPU conv kernels, BatchNorm / loss restated) is timed instead."""
    import types

    import torch
    from oracle import vae_oracle as O
    cfg = O.VAEConfig(input_image_size=64, z_dimension=64)
    st = O.init_state(cfg, seed=0)
    x = O.normalise(O.synthetic_labels(batch, 64))
    if os.path.exists(os.path.join(REF_DIR, "model.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_reference_model", os.path.join(REF_DIR, "model.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        m = ref.VAE(in_channels=1, intermediate_channels=32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64,
                    pixelcnn=False, only_pixelcnn=False, nll=1, kl=1, mmd=0, require_rsample=True, sigma_decoder=0.1,
                    input_image_size=64)
        m.load_state_dict(st, strict=True)
        m.train(True)
        params = list(m.parameters())
        dev, largs = torch.device("cpu"), types.SimpleNamespace(data_ratio_of_labels=None)

        def step():
            mu, logvar, enc, recon = m(x)                                            # main.py:389
            loss, _, _, _ = m.loss(x, mu, logvar, enc, recon, dev, largs)            # main.py:390
            for p in params:
                p.grad = None                                                        # main.py:397
            loss.backward()                                                          # main.py:398
            return float(loss.detach())
        return step, "reference"
    eps = torch.randn(batch, 64, 1, 1, generator=torch.Generator().manual_seed(4321))
    timed = O.make_timed_step(st, cfg)
    return (lambda: timed(x, x, eps)), "port"


def cpu_reference_fps(steps, warmup, batch=32):
    """The reference's train step (forward -> loss -> backward, main.py:389-390,398) on the host cores, all of them."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = _reference_step(batch)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med * 1e3, cores, torch.get_num_threads(), kind


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path on this box's host cores, on the SAME workload as
    our arm (BASELINE configs[1]: 256 frames per step -- BatchNorm statistics over the same batch), a bounded number of steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(3, min(args.steps, 20))
    warm = min(max(args.warmup, 1), 2)
    n = args.batch
    fps, ms, cores, threads, kind = cpu_reference_fps(steps, warm, batch=n)
    line = {
        "impl": "reference", "metric": "train frames/sec (fwd+bwd)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "model.py VAE z=64 64x64 Gaussian-NLL sigma=0.1, fwd+loss+bwd (BASELINE configs[1])",
                   "frames_per_gpu": n, "global_batch": n, "seq_len": 20, "parallelism": "cpu", "precision": "fp32",
                   "launch": f"host threads ({threads})"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{steps} steps of {n} frames on {cores} host cores (median), fp32, "
                                   + ("the reference's own model.py VAE" if kind == "reference" else "oracle port of model.py")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


NB_TRAIN_MFLOP_PER_FRAME = 7714.586624     # notebook variant at 128x128 (mmvae_layout.train_flops)


def nb_cpu_fps(steps, warmup, batch=4, size=128):
    """The notebook's loop body (vae-kl.ipynb:210-233, forward + CE/KL + backward) on the host cores: the notebook's own
    VAE_Encoder / VAE_Decoder classes (code cell 5 of oracle/_ref/vae-kl.ipynb, executed unmodified) when the copy is there
    (kind 'reference'), else the oracle port."""
    import torch
    from oracle import nb_oracle as NB
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = NB.NbConfig(image_size=size)
    st = NB.init_state(cfg, seed=0)
    x, y = NB.synthetic_batch(cfg, batch, seed=1234)
    eps = torch.randn(batch, cfg.z_dimensions, cfg.latent_hw, cfg.latent_hw, generator=torch.Generator().manual_seed(4321))
    nb_path = os.path.join(REF_DIR, "vae-kl.ipynb")
    kind = "port"
    if os.path.exists(nb_path):
        import torch.nn.functional as F
        from torch.distributions import Normal, kl_divergence
        cells = [c for c in json.load(open(nb_path))["cells"] if c["cell_type"] == "code"]
        ns = {}
        exec("import torch\nfrom torch import nn\nfrom torch.nn import functional as F\n"
             "from torch.distributions import Normal\n" + "".join(cells[5]["source"]), ns)
        enc = ns["VAE_Encoder"](cfg.in_channels, cfg.channels, cfg.z_dimensions)
        dec = ns["VAE_Decoder"](cfg.in_channels, cfg.channels, cfg.z_dimensions)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in st.items() if k.startswith("encoder.")}, strict=True)
        dec.load_state_dict({k[len("decoder."):]: v for k, v in st.items() if k.startswith("decoder.")}, strict=True)
        params = list(enc.parameters()) + list(dec.parameters())
        kind = "reference"

        def step(x, y, eps):
            mu, logvar = enc(x)                                                    # vae-kl.ipynb:213
            encoding = enc.rsample(mu, logvar)
            recon = dec(encoding)
            pxz = (F.cross_entropy(recon, y, reduction="none") / x.shape[0]).sum() # vae-kl.ipynb:225
            kl = (kl_divergence(Normal(mu, (0.5 * logvar).exp()), Normal(torch.tensor(0.), torch.tensor(1.))) / x.shape[0]).sum()
            loss = pxz + kl
            for p in params:
                p.grad = None
            loss.backward()
            return float(loss.detach())
    else:
        step = NB.make_timed_step(st, cfg)
    for _ in range(warmup):
        step(x, y, eps)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step(x, y, eps)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med * 1e3, cores, torch.get_num_threads(), kind


def run_reference_notebook(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warm = min(args.steps, 5), min(args.warmup, 1)
    fps, ms, cores, threads, kind = nb_cpu_fps(steps, warm)
    emit({
        "impl": "reference", "metric": "train frames/sec (fwd+bwd)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vae-kl.ipynb VAE 128x128, CE over 256 grey levels + KL, fwd+loss+bwd (BASELINE configs[4])",
                   "batch_per_step": 4},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{steps} steps of 4 frames of 128x128 on {cores} host cores (median), fp32"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


class Ctx:
    """Process-group plumbing shared by the workloads: one process per GPU, barrier + synchronize on both sides of every
    timed block, device time (CUDA events) as the MAX over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed_blocks(self, fn, steps, blocks, finish=None):
        """`blocks` timed blocks of EXACTLY `steps` steps each; returns the per-block milliseconds (max over ranks)."""
        torch = self.torch
        out = []
        it = 0
        for _ in range(blocks):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            ev0.record()
            for _ in range(steps):
                fn(it)
                it += 1
            if finish is not None:
                finish()
            ev1.record()
            self.barrier()
            ms = ev0.elapsed_time(ev1)
            if self.world > 1:
                t = torch.tensor([ms], device=self.dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                ms = float(t.item())
            out.append(ms)
        return out

    def ranks_agree(self, flat):
        """Is `flat` (the gradient arena after the exchange) bit-identical on every rank?"""
        if self.world == 1:
            return True
        torch = self.torch
        parts = [torch.empty_like(flat) for _ in range(self.world)]
        self.dist.all_gather(parts, flat.contiguous())
        return all(torch.equal(parts[0], q) for q in parts[1:])

    def close(self):
        if self.world > 1:
            # never let a stuck teardown outlive the measurement
            watchdog = threading.Timer(30.0, lambda: os._exit(0))
            watchdog.daemon = True
            watchdog.start()
            self.torch.cuda.synchronize()
            self.dist.barrier()
            self.dist.destroy_process_group()
            watchdog.cancel()


def median(v):
    s = sorted(v)
    return s[len(s) // 2]


def load_traffic():
    """dram read + write bytes per launch of the kernels the rooflines quote, from the committed ncu --set full summary
    (profiles/r02_traffic.csv: kernel key, dram__bytes_read.sum, dram__bytes_write.sum).  {} when absent."""
    out = {}
    path = os.path.join(ROOT, "profiles", "r02_traffic.csv")
    if os.path.exists(path):
        for row in open(path).read().splitlines():
            if not row or row.startswith("#") or row.startswith("key,"):
                continue
            f = row.split(",")
            try:
                out[f[0]] = int(float(f[1]) + float(f[2]))
            except (ValueError, IndexError):
                pass
    return out


def run_notebook(args, ctx, n, full=True):
    """BASELINE configs[4]: the notebook variant (vae-kl.ipynb) on 128x128 frames, 512 frames per GPU, bf16, weak scaling."""
    import ctypes
    import gc

    torch = ctx.torch
    import mmvae_b200 as M
    from mmvae_b200 import parallel as PAR

    rank, world, dev, size = ctx.rank, ctx.world, ctx.dev, 128
    steps = args.steps if full else max(5, min(args.steps, 40))
    blocks = args.blocks if full else 3
    torch.manual_seed(0)
    model = M.NotebookVAE(1, 32, 32, image_size=size, precision=args.precision).to(dev)
    if world > 1:
        PAR.data_parallel(model)

    def frames(seed):
        """20-frame sequences of two bouncing grey blobs (56x56 on 128x128), flattened to frames, uint8 grey levels"""
        g = torch.Generator().manual_seed(seed)
        d, nseq = 56, (n + 19) // 20
        out = torch.zeros(nseq * 20, size, size, dtype=torch.uint8)
        for s in range(nseq):
            for _ in range(2):
                blob = (torch.rand(d, d, generator=g) * 255).to(torch.uint8)
                pos = torch.rand(2, generator=g) * (size - d)
                vel = (torch.rand(2, generator=g) - 0.5) * 12
                for t in range(20):
                    oy, ox = int(pos[0]), int(pos[1])
                    out[s * 20 + t, oy:oy + d, ox:ox + d] = torch.maximum(out[s * 20 + t, oy:oy + d, ox:ox + d], blob)
                    pos = pos + vel
                    for k in range(2):
                        if pos[k] < 0 or pos[k] > size - d:
                            vel[k] = -vel[k]
                            pos[k] = pos[k].clamp(0, size - d)
        return out[:n].contiguous()

    n_batches = 2
    host = [frames(1234 + 97 * rank + b).pin_memory() for b in range(n_batches)]
    resident = [model.prepare_input(h.to(dev)) for h in host]
    info = M._lib.layout(model._desc(n, True))

    def step_resident(i):
        x, y = resident[i % n_batches]
        return model.train_step(x, y)

    reader = LossReader(torch)

    from mmvae_b200.data import Prefetcher
    pre = Prefetcher(dev)
    pre.push(host[0])                                             # H2D of the first batch

    def step_e2e(i):
        f = pre.pop()                                             # the staged batch (the stream waits for its copy)
        x, y = model.prepare_input(f)                             # normalisation + int64 targets on the device
        loss = model.train_step(x, y)[0]
        pre.push(host[(i + 1) % n_batches])                       # H2D of the NEXT batch from pinned memory, on a copy stream,
                                                                  # beside the step just enqueued
        reader.push(i, loss)                                      # D2H of the loss into pinned memory, read one step late

    for i in range(max(3, args.warmup if full else 3)):
        step_resident(i)
    with ClockSampler(ctx.local) as clk:
        l0 = M._lib.lib.mmvae_launch_count()
        ms_blocks = ctx.timed_blocks(step_resident, steps, blocks)
        launches = (M._lib.lib.mmvae_launch_count() - l0) // blocks
    ms = median(ms_blocks)
    agree = ctx.ranks_agree(model.flat_grads)
    for i in range(2):
        step_e2e(i)
    reader.finish()
    e2e_blocks = ctx.timed_blocks(step_e2e, steps, blocks, finish=reader.finish)
    ms_e2e = median(e2e_blocks)
    assert reader.reads == steps * blocks + 2 and reader.total == reader.total, "every step's loss must have been read (and be finite)"

    sustained, burst, hbm, which = peaks()
    roofs = []
    if full and rank == 0 and args.precision == "bf16":
        # the three dedicated tcgen05 kernels of decoder.conv4 (94 % of the step's FLOPs), each timed alone on the tensors the
        # last step left in the workspace; inputs (4.8 GB) are far larger than L2
        desc, ws, x_last = model._state
        y_last = resident[(steps * blocks - 1) % n_batches][1]
        scratch = torch.zeros(model._n_params, dtype=torch.float32, device=dev)
        ab, af = ctypes.c_int64(), ctypes.c_int64()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        names = ["nb_tail_fwd_kernel: decoder.conv4 forward (32->256, 3x3, 128x128) fused with softmax cross-entropy, writes d logits",
                 "nb_tail_dgrad_kernel: decoder.conv4 data gradient (transposed form, col2im in TMEM + epilogue)",
                 "nb_tail_wgrad_kernel: decoder.conv4 weight + bias gradient (pixel axis as K)"]
        tr = load_traffic()
        traffic = [tr.get("nb_tail_fwd_kernel"), tr.get("nb_tail_dgrad_kernel"), tr.get("nb_tail_wgrad_kernel")]
        for which_k in (0, 2, 1):                                 # d logits must exist before the gradients read them
            def one():
                M._lib.check(M._lib.lib.mmvae_nb_bench_tail(ctypes.byref(desc), which_k, ctypes.c_void_p(model._arena.data_ptr()),
                                                            ctypes.c_void_p(y_last.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                                            ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(ab), ctypes.byref(af), stream),
                             "mmvae_nb_bench_tail")
            for _ in range(2):
                one()
            evs = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); one(); e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            us = sum(a.elapsed_time(b) for a, b in evs) / len(evs) * 1e3
            tfs, gbs = af.value / (us * 1e-6) / 1e12, ab.value / (us * 1e-6) / 1e9
            roofs.append({"bound": "tensor", "unit": "TFLOP/s", "achieved": tfs, "peak": burst, "frac": tfs / burst,
                          "traffic": traffic[which_k], "kernel": names[which_k], "algorithmic_bytes_per_launch": ab.value,
                          "flops_per_launch": af.value, "us_per_launch": us, "gbs": gbs, "hbm_frac": gbs / hbm,
                          "note": f"of the {which} burst bf16 peak (kernel timed alone); hbm_frac = algorithmic GB/s over the {which} HBM copy peak; "
                                  "traffic = dram read+write of one launch from profiles/r02_traffic.csv (ncu --set full)"
                                  + ("; the forward kernel is bound by its 4.26 GB write stream: a pure fill of that size runs at 3.9 TB/s "
                                     "(scripts/hbm_write_probe.py)" if which_k == 0 else "")})
        roofs.sort(key=lambda r: -r["us_per_launch"])

    fps = world * n * steps / (ms * 1e-3)
    fps_e2e = world * n * steps / (ms_e2e * 1e-3)
    tf = fps / world * NB_TRAIN_MFLOP_PER_FRAME * 1e6 / 1e12
    line = {
        "metric": "train frames/sec (fwd+bwd)", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "vae-kl.ipynb VAE 128x128, CE over 256 grey levels + KL, fwd+loss+bwd (BASELINE configs[4])",
                   "frames_per_gpu": n, "global_batch": n * world, "seq_len": 20, "parallelism": f"dp{world}",
                   "precision": args.precision, "launch": f"host ({launches // max(steps, 1)} launches per step)",
                   "l2": f"activation workspace {info.workspace_bytes / 1e9:.1f} GB streamed every step (>> 126 MB L2), "
                         f"{n_batches} rotating input batches",
                   "timing": f"median of {blocks} blocks of {steps} steps, each block bracketed by barrier + synchronize, CUDA events, max over ranks"},
        "ms_per_step_blocks": [b / steps for b in ms_blocks],
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": n * size * size, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / steps,
                "mode": "per step: H2D of the uint8 frames from pinned memory (double-buffered on a copy stream: the copy of batch i+1 runs beside step i, mmvae_b200.data.Prefetcher), device normalisation, train step, D2H of the loss into "
                        "pinned memory; the host reads each loss one step late (asynchronous logging), all inside the timed region"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "roofline": roofs[0] if roofs else {"bound": "tensor", "achieved": tf, "peak": sustained, "unit": "TFLOP/s",
                                            "frac": tf / sustained, "traffic": None},
        "roofline_others": roofs[1:],
        "step_tensor": {"achieved": tf, "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained,
                        "note": f"whole step, 7714.59 MFLOP/frame algorithmic, per GPU, of {which} sustained bf16 peak"},
    }
    if world > 1:
        line["grads_bit_identical_across_ranks"] = bool(agree)
        assert agree, "the gradient arena differs between ranks after the all-reduce"
    if full and rank == 0 and world == 1 and not args.no_cpu_baseline:
        cfps, cms, cores, threads, kind = nb_cpu_fps(3, 1)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": kind,
                                "sample": f"3 steps of 4 frames of 128x128 on {cores} host cores, fp32, median"}
    del model, resident, pre
    gc.collect()
    torch.cuda.empty_cache()
    return line


class LossReader:
    """Device -> host read of every step's loss without stalling the stream: the scalar is copied into one of two pinned
    host slots right behind the step (D2H inside the timed region) and the host reads it one step later, after the slot's
    event -- what a training loop that logs asynchronously does.  finish() reads the last one."""

    def __init__(self, torch):
        self.torch = torch
        self.slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.pending = None
        self.total = 0.0
        self.reads = 0

    def push(self, i, loss_dev):
        k = i & 1
        if self.pending is not None and self.pending == k:      # never overwrite a slot that was not read yet
            self._read(self.pending)
        self.slots[k].copy_(loss_dev.detach().reshape(1), non_blocking=True)
        self.events[k].record()
        prev = self.pending
        self.pending = k
        if prev is not None and prev != k:
            self._read(prev)

    def _read(self, k):
        self.events[k].synchronize()
        self.total += float(self.slots[k][0])
        self.reads += 1

    def finish(self):
        if self.pending is not None:
            self._read(self.pending)
            self.pending = None


_REAL_STDOUT = None


def guard_stdout():
    """Libraries (NCCL's version banner, for one) write to file descriptor 1; the contract is ONE JSON line on stdout.
    Keep the real stdout aside for that line and send everything else written to fd 1 to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()



def run_resnet(args, ctx, width, zdim, n, full=True):
    """BASELINE configs[1] (width 1, the headline) / configs[3] (width 2, z = 256): the model.py VAE, bf16, weak scaling."""
    import gc
    import types

    torch = ctx.torch
    import mmvae_b200 as M
    from mmvae_b200 import data as D
    from mmvae_b200 import parallel as PAR

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    steps = args.steps if full else max(5, min(args.steps, 100))
    blocks = args.blocks if full else 3
    torch.manual_seed(0)
    model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=zdim, pixelcnn=False,
                  only_pixelcnn=False, nll=1, kl=1, mmd=0, sigma_decoder=0.1, input_image_size=64,
                  precision=args.precision, width=width).to(dev).train()
    model.defer_metrics = True
    if world > 1:
        PAR.data_parallel(model)
    largs = types.SimpleNamespace(data_ratio_of_labels=None)

    # several distinct batches per rank; the step's working set (activation workspace, ~0.5 GB at N=256)
    # is streamed through once per step and is several times the 126 MB L2
    n_batches = 4
    labels_host = [D.synthetic_labels(n, 64, seed=1234 + 97 * rank + b).pin_memory() for b in range(n_batches)]
    x_dev = [D.prepare_input(l.to(dev)) for l in labels_host]
    ws_bytes = M._lib.layout(model._desc(n, True)).workspace_bytes
    params = list(model.parameters())
    reader = LossReader(torch)
    gstep = gstep_lab = gstep_adam = None
    launches_per_step = None

    if args.no_graph:
        def step_resident(i):
            x = x_dev[i % n_batches]
            mu, logvar, enc, recon = model(x)
            loss, pxz, kl, _ = model.loss(x, mu, logvar, enc, recon, dev, largs)
            for p in params:
                p.grad = None                      # optimizer.zero_grad(), main.py:397
            loss.backward()
            return loss

        def step_e2e(i):
            lab = labels_host[i % n_batches].to(dev, non_blocking=True)     # H2D from pinned memory
            x = D.prepare_input(lab)                                        # main.py:383-388 on the device
            mu, logvar, enc, recon = model(x)
            loss, pxz, kl, _ = model.loss(x, mu, logvar, enc, recon, dev, largs)
            for p in params:
                p.grad = None
            loss.backward()
            reader.push(i, loss)                                            # D2H of the loss into pinned memory, read one step late
    else:
        # the same loop body captured once (through VAE.forward / VAE.loss / backward) and replayed: mmvae_b200/graph.py
        l0 = M._lib.lib.mmvae_launch_count()
        gstep = M.GraphedTrainStep(model, n, args=largs, warmup=1)
        launches_per_step = (M._lib.lib.mmvae_launch_count() - l0) // 2       # 1 warm-up + 1 captured pass
        gstep_lab = M.GraphedTrainStep(model, n, args=largs, warmup=1, from_labels=(D.DATA_MEAN, D.DATA_STD))

        def step_resident(i):
            gstep.x.copy_(x_dev[i % n_batches])                               # device-resident batch -> static input
            return gstep(None)[0]

        gstep_lab.prefetch(labels_host[0])                                    # H2D of the first batch

        def step_e2e(i):
            loss = gstep_lab()[0]                                             # staged batch -> static input, then the graph
            gstep_lab.prefetch(labels_host[(i + 1) % n_batches])              # H2D of the NEXT batch from pinned memory, on a
                                                                              # copy stream, beside the replay just launched
            reader.push(i, loss)                                              # D2H of the loss into pinned memory, read one step late

    for i in range(args.warmup if full else 3):
        step_resident(i)
    with ClockSampler(ctx.local) as clk:
        l0 = M._lib.lib.mmvae_launch_count()
        ms_blocks = ctx.timed_blocks(step_resident, steps, blocks)
        launches = (M._lib.lib.mmvae_launch_count() - l0) // blocks
    ms = median(ms_blocks)
    if launches_per_step is not None:
        launches = launches_per_step * steps                                  # replayed kernels of the captured step
    agree = ctx.ranks_agree(model.last_flat_grad)
    for i in range(3):
        step_e2e(i)
    reader.finish()
    e2e_blocks = ctx.timed_blocks(step_e2e, steps, blocks, finish=reader.finish)
    ms_e2e = median(e2e_blocks)
    assert reader.reads == steps * blocks + 3 and reader.total == reader.total, "every step's loss must have been read (and be finite)"

    # ---- the whole loop body main.py:389-399: the step above + optimizer.step() (mmvae_b200.FusedAdam, captured in the graph)
    with_adam = None
    if full and not args.no_graph:
        saved = model.flat_parameters.clone()
        opt = M.FusedAdam(model, lr=1e-3)
        gstep_adam = M.GraphedTrainStep(model, n, args=largs, warmup=1, optimizer=opt)

        def step_adam(i):
            gstep_adam.x.copy_(x_dev[i % n_batches])
            return gstep_adam(None)[0]
        for i in range(3):
            step_adam(i)
        adam_blocks = ctx.timed_blocks(step_adam, steps, min(blocks, 3))
        ms_adam = median(adam_blocks)
        loss_last = float(gstep_adam.loss)
        with_adam = {"value": world * n * steps / (ms_adam * 1e-3), "unit": "frames/s", "ms_per_step": ms_adam / steps,
                     "loss_after": loss_last, "finite": bool(torch.isfinite(model.flat_parameters).all()),
                     "replicas_bit_identical": bool(ctx.ranks_agree(model.flat_parameters)),
                     "note": "forward + loss + zero_grad + backward + optimizer.step() (main.py:389-399) as one graph replay; "
                             "mmvae_b200.FusedAdam, one kernel over the flat fp32 arena, lr 1e-3"}
        with torch.no_grad():
            model.flat_parameters.copy_(saved)

    # ---- rooflines, measured live: one launch per iteration of the production kernel of a (conv, direction) pair through
    # mmvae_bench_conv on the tensors the last step left in the workspace, CUDA events around each launch on the launching
    # stream, L2 flushed (256 MB memset) between iterations.  Algorithmic bytes = bf16 input + output + weights, each
    # touched once; `traffic` = dram read + write of the same launch from the ncu --set full capture in profiles/.
    #   roofline         the kernel with the largest share of the step (gconv_tc_kernel, the tcgen05 implicit-GEMM conv)
    #                    on its most expensive launch, encoder.layer4.0.conv2 (256->256, 3x3 on 2x2 maps: M = 1024,
    #                    N = 256, K = 2304); its arithmetic intensity (542 FLOP/B) is above the ridge: tensor pipe
    #   roofline_others  the shared-memory-resident band kernels (slab_tc.cu) on the heaviest-traffic layer of the model,
    #                    decoder.uplayer5.0.conv2 (16->16 transposed conv, 256x32x32 -> 256x64x64): forward, data
    #                    gradient, weight gradient -- HBM-bound
    roof, roof_others = None, []
    if full:
        step_resident(0)                       # every rank (the replay holds the exchange): leaves a step's tensors in the workspace
        torch.cuda.synchronize()
    if full and rank == 0 and args.precision == "bf16" and width == 1:    # per-kernel rooflines: the headline configuration only
        import ctypes
        desc, ws, _info = model._workspace(n, True)
        names = [c[0] for c in M._lib.conv_table(desc)]
        ab, af = ctypes.c_int64(), ctypes.c_int64()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        scratch = torch.zeros(model._n_params, dtype=torch.float32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        sustained_tf, burst_tf, hbm_peak, which_peak = peaks()
        tr = load_traffic()

        def measure(conv, direction):
            ci = names.index(conv)

            def one():
                M._lib.check(M._lib.lib.mmvae_bench_conv(ctypes.byref(desc), ci, direction, ctypes.c_void_p(model._arena.data_ptr()),
                                                         ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                                         ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(ab), ctypes.byref(af),
                                                         stream), "mmvae_bench_conv")
            for _ in range(3):
                one()
            evs = []
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); one(); e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            us = [a.elapsed_time(b) * 1e3 for a, b in evs]
            return sum(us) / len(us), ab.value, af.value

        def entry(kernel, key, conv, direction, bound):
            k_us, nbytes, nflops = measure(conv, direction)
            gbs, tfs = nbytes / (k_us * 1e-6) / 1e9, nflops / (k_us * 1e-6) / 1e12
            e = {"bound": bound, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                 "achieved": gbs if bound == "hbm" else tfs, "peak": hbm_peak if bound == "hbm" else burst_tf,
                 "traffic": tr.get(key), "kernel": kernel, "algorithmic_bytes_per_launch": nbytes, "flops_per_launch": nflops,
                 "us_per_launch": k_us, "gbs": gbs, "tensor_tflops": tfs,
                 "note": f"of the {which_peak} " + ("HBM copy peak" if bound == "hbm" else "burst bf16 peak (kernel timed alone)") +
                         "; traffic = dram read+write of one launch from profiles/r02_traffic.csv (ncu --set full)"}
            e["frac"] = e["achieved"] / e["peak"]
            return e

        if "encoder.layer4.0.conv2" in names and "decoder.uplayer5.0.conv2" in names:
            roof = entry("gconv_tc_kernel forward on encoder.layer4.0.conv2 (Conv2d 256->256 k3 s1 p1, 256x2x2): M=1024 N=256 K=2304",
                         "gconv_tc_kernel:encoder.layer4.0.conv2:fwd", "encoder.layer4.0.conv2", 0, "tensor")
            u5 = "decoder.uplayer5.0.conv2 (ConvTranspose2d 16->16 k4 s2 p1, 256x32x32 -> 256x64x64)"
            roof_others = [entry("slab_fwd_kernel<16> forward on " + u5, "slab_fwd_kernel:decoder.uplayer5.0.conv2", "decoder.uplayer5.0.conv2", 0, "hbm"),
                           entry("slab_dgrad_kernel<16> data gradient on " + u5, "slab_dgrad_kernel:decoder.uplayer5.0.conv2", "decoder.uplayer5.0.conv2", 1, "hbm"),
                           entry("slab_wgrad_kernel<16> weight gradient on " + u5, "slab_wgrad_kernel:decoder.uplayer5.0.conv2", "decoder.uplayer5.0.conv2", 2, "hbm")]
        del flush, scratch

    fps = world * n * steps / (ms * 1e-3)
    fps_e2e = world * n * steps / (ms_e2e * 1e-3)
    sustained, burst, hbm, which = peaks()
    mflop_per_frame = M._lib.layout(model._desc(n, True)).train_flops / n / 1e6      # 227.02 base, 896.01 widened
    tflops_per_gpu = fps / world * mflop_per_frame * 1e6 / 1e12
    mb_per_frame = 2.790 if width == 1 else 5.550                                      # SURVEY.md 8(d), bf16 algorithmic bytes
    step_gbs = fps / world * mb_per_frame * 1e6 / 1e9
    line = {
        "metric": "train frames/sec (fwd+bwd)", "value": fps, "unit": "frames/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": ("model.py VAE z=64 64x64 Gaussian-NLL sigma=0.1, fwd+loss+bwd (BASELINE configs[1])" if width == 1 else
                                "model.py VAE widened 2x channels z=256 64x64 Gaussian-NLL sigma=0.1, fwd+loss+bwd (BASELINE configs[3])"),
                   "frames_per_gpu": n, "global_batch": n * world, "seq_len": 20,
                   "parallelism": f"dp{world}", "precision": args.precision,
                   "launch": "host" if args.no_graph else "cuda-graph replay of the captured step",
                   "l2": f"activation workspace {ws_bytes / 1e6:.0f} MB streamed every step (> 126 MB L2), "
                         f"{n_batches} rotating input batches",
                   "timing": f"median of {blocks} blocks of {steps} steps, each block bracketed by barrier + synchronize, CUDA events, max over ranks",
                   "step": "forward + loss (incl. the MMD diagnostic, model.py:394-396) + zero_grad + backward; optimizer.step() is reported in with_adam"},
        "ms_per_step_blocks": [b / steps for b in ms_blocks],
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": n * 64 * 64, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / steps,
                "mode": "per step: H2D of the uint8 label maps from pinned memory (double-buffered on a copy stream: the copy of batch i+1 runs beside step i, GraphedTrainStep.prefetch), graph replay (normalisation + train step), D2H of "
                        "the loss into pinned memory; the host reads each loss one step late (asynchronous logging), all inside the "
                        "timed region"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof if roof is not None else {"bound": "tensor", "achieved": tflops_per_gpu, "peak": sustained,
                                                   "unit": "TFLOP/s", "frac": tflops_per_gpu / sustained, "traffic": None},
        "roofline_others": roof_others,
        "step_tensor": {"achieved": tflops_per_gpu, "peak": sustained, "unit": "TFLOP/s", "frac": tflops_per_gpu / sustained,
                        "note": f"whole step, {mflop_per_frame:.2f} MFLOP/frame algorithmic, per GPU, of {which} sustained bf16 peak"},
        "step_hbm": {"achieved": step_gbs, "peak": hbm, "unit": "GB/s", "frac": step_gbs / hbm,
                     "note": f"whole step, {mb_per_frame} MB/frame algorithmic bf16 activation traffic (SURVEY.md 8(d)), per GPU, of the {which} HBM copy peak"},
    }
    if launches_per_step is not None:
        line["launches_per_step"] = int(launches_per_step)
    if with_adam is not None:
        line["with_adam"] = with_adam
    if world > 1:
        line["grads_bit_identical_across_ranks"] = bool(agree)
        assert agree, "the gradient arena differs between ranks after the all-reduce"
    if full and rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = min(n, 256) if width == 1 else 32
        cfps, cms, cores, threads, kind = cpu_reference_fps(8, 2, batch=cb)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": kind,
                                "sample": f"8 steps of {cb} frames (the same workload, base width) on {cores} host cores, fp32, median, "
                                          + ("the reference's own model.py VAE" if kind == "reference" else "oracle port of model.py")}
    # the captured graphs hold NCCL work on the communicator: release them before anything tears it down
    for g in (gstep, gstep_lab, gstep_adam):
        if g is not None:
            g.graph = None
    del step_resident, step_e2e, gstep, gstep_lab, gstep_adam, model
    gc.collect()
    torch.cuda.empty_cache()
    return line


def brief(line):
    """What an extra workload contributes to the headline JSON line."""
    keep = ("value", "unit", "ms_per_step", "ms_per_step_blocks", "steps", "n_gpus", "dtype", "gpu_launches", "launches_per_step",
            "clocks", "step_tensor", "step_hbm", "grads_bit_identical_across_ranks")
    out = {k: line[k] for k in keep if k in line}
    out["workload"] = line["config"]["workload"]
    out["frames_per_gpu"] = line["config"]["frames_per_gpu"]
    out["e2e"] = {k: line["e2e"][k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    return out


def main():
    guard_stdout()
    # a stalled collective must not hold the box until the caller's limit: dump every thread's stack and leave
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MMVAE_BENCH_STALL_S", "900")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--blocks", type=int, default=5, help="timed blocks of --steps steps each; the median block is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU (default 256; 512 for --workload notebook)")
    ap.add_argument("--workload", default="resnet", choices=["resnet", "widened", "notebook"],
                    help="resnet: model.py VAE, BASELINE configs[1] (the headline); widened: 2x channels, z=256, 128 frames per GPU, "
                         "configs[3]; notebook: vae-kl.ipynb VAE on 128x128, configs[4]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no `workloads` block for configs[3] / configs[4])")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {"notebook": 512, "widened": 128}.get(args.workload, PER_GPU_BATCH)
    if args.impl == "reference":
        (run_reference_notebook if args.workload == "notebook" else run_reference)(args)
        return
    args.warmup = max(args.warmup, 3)
    args.blocks = max(args.blocks, 1)
    ctx = Ctx()
    if args.workload == "notebook":
        line = run_notebook(args, ctx, args.batch)
    elif args.workload == "widened":
        line = run_resnet(args, ctx, 2, 256, args.batch)
    else:
        line = run_resnet(args, ctx, 1, 64, args.batch)
        if not args.no_extra and args.precision == "bf16":
            # the other configurations BASELINE.json names, in the same process and on the same ranks, so that every claimed
            # configuration has a record in the one JSON line: configs[3] (widened, 128 frames per GPU) and configs[4]
            # (notebook variant, 512 frames of 128x128 per GPU).  Same timing rules, fewer steps.
            extras = {}
            for name, fn in (("widened", lambda: run_resnet(args, ctx, 2, 256, 128, full=False)),
                             ("notebook", lambda: run_notebook(args, ctx, 512, full=False))):
                try:
                    extras[name] = brief(fn())
                except Exception as e:                                       # the headline must survive an extra's failure
                    extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            line["workloads"] = extras
    if ctx.rank == 0:
        emit(line)
    ctx.close()


if __name__ == "__main__":
    main()
