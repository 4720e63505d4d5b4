"""GPU: the tcgen05 kernels against the fp32-FMA SIMT kernels on identical inputs, conv by conv
(mmvae_selftest_tc, include/mmvae.h): forward output + BatchNorm statistics, weight gradient and data
gradient of every conv the tensor-core path covers.  Weights are bf16-representable so both paths multiply
the same numbers; what remains is fp32 summation order and 1-ulp bf16 rounding flips of the stored outputs."""
from ctypes import byref, c_void_p

import pytest
import torch

import mmvae_b200 as M
from mmvae_b200 import _lib

pytestmark = pytest.mark.gpu


def run_selftest(n, image_size=64, z=64, width=1, out_channels=1):
    desc = _lib.make_desc(n, 1, out_channels, z, image_size, width=width)
    info = _lib.layout(desc)
    convs = _lib.conv_table(desc)
    gen = torch.Generator().manual_seed(7)
    params = (0.05 * torch.randn(info.n_params, generator=gen)).bfloat16().float().cuda()
    ws = torch.zeros(info.workspace_bytes, dtype=torch.uint8, device="cuda")
    ga = torch.zeros(info.n_params, device="cuda")
    gb = torch.zeros(info.n_params, device="cuda")
    rep = torch.zeros(len(convs) * 16, device="cuda")
    s = c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib.mmvae_selftest_tc(byref(desc), c_void_p(params.data_ptr()), c_void_p(ws.data_ptr()), ws.numel(),
                                          c_void_p(ga.data_ptr()), c_void_p(gb.data_ptr()), c_void_p(rep.data_ptr()),
                                          rep.numel(), s), "mmvae_selftest_tc")
    torch.cuda.synchronize()
    rep = rep.view(len(convs), 4, 4).cpu().double()
    rows = []
    for (name, *_), r in zip(convs, rep):
        rel = [(r[j, 0].sqrt() / r[j, 1].sqrt().clamp_min(1e-30)).item() if r[j, 1] > 0 else None for j in range(4)]
        rows.append((name, rel))
    return rows


# (256, 64) is the bench batch: persistent CTAs walk several tiles / bands each (ring wrap-around, both TMEM buffers)
@pytest.mark.parametrize("n,size", [(8, 64), (64, 64), (256, 64), (6, 28), (3, 32), (160, 32)])
def test_tc_kernels_match_simt(n, size):
    rows = run_selftest(n, image_size=size, z=64 if size == 64 else 32)
    bad = []
    tested = 0
    for name, rel in rows:
        print(f"{name:36s} " + " ".join("   --   " if r is None else f"{r:8.1e}" for r in rel))
        tol = (4e-3, 1e-3, 1e-3, 4e-3)          # forward y (bf16), BN (mean, rstd), dW (fp32), dX (bf16)
        for j, r in enumerate(rel):
            if r is None:
                continue
            tested += 1
            if not r <= tol[j]:
                bad.append((name, ("fwd", "bn", "wgrad", "dgrad")[j], r))
    assert tested >= 80
    assert not bad, bad


def test_tc_kernels_match_simt_widened():
    """BASELINE configs[3]: 2x channels, latent dim 256 (64-channel stem, 512-channel bottleneck, 128-column tiles)."""
    rows = run_selftest(8, image_size=64, z=256, width=2)
    tol = (4e-3, 1e-3, 1e-3, 4e-3)
    bad = [(name, j, r) for name, rel in rows for j, r in enumerate(rel) if r is not None and not r <= tol[j]]
    assert not bad, bad


def test_tc_kernels_match_simt_cluster_multicast():
    """The optional deep-K path of gconv_tc_kernel (off by default, MMVAE_MC_MIN_CHUNKS): the channel tiles of a pixel tile
    form a thread-block cluster and multicast their slice of every A box (UTMALDG.MULTICAST, multicast commit, drain lap).
    The switch is read once per process, so the check runs in a child process."""
    import os, subprocess, sys
    code = ("import sys; sys.path.insert(0, 'tests'); import test_gpu_tc_selftest as t; "
            "rows = t.run_selftest(64, image_size=64, z=64); tol = (4e-3, 1e-3, 1e-3, 4e-3); "
            "bad = [(n, j, r) for n, rel in rows for j, r in enumerate(rel) if r is not None and not r <= tol[j]]; "
            "assert not bad, bad; print('ok', len(rows))")
    env = dict(os.environ, MMVAE_MC_MIN_CHUNKS="4")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout
