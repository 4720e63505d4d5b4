"""GPU: parity of the CUDA path (through the drop-in module API -> C ABI) against the golden fixtures made
from the live reference and against the CPU oracle.

Tolerances (relative L2 per tensor):
  * fp32 validation mode: north_star asks 1e-5, but the reference's OWN fp32 result sits 2.5e-5 (N=4) to
    1.8e-4 (N=32) away from an fp64 evaluation of the same formulas (tests/test_oracle_golden.py),
    so the test is two-sided: (a) against the reference fixture within 5e-4 (its own noise floor), and
    (b) against the fp64 oracle no worse than 3x the reference-fp32's distance from fp64, floor 1e-5.
  * bf16 mode: 1e-2 on the loss; gradients are bounded by the intrinsic bf16-storage floor measured with the
    oracle's emulate_bf16 mode (see test_bf16_step_within_bf16_floor).
"""
import numpy as np
import pytest
import torch

from golden_util import CASES, Golden, sample_idx
from oracle import vae_oracle as O
from ours_util import build_model, rel_l2, train_step, workspace_tensor

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_fp32_step_matches_reference_fixture(name):
    g = Golden(name)
    m = build_model(g.cfg, g.state(), "fp32")
    res = train_step(m, g.cfg, g.x, g.target, g.eps, g.ce_weight)
    g.check_step(res, rtol=5e-4)


@pytest.mark.parametrize("name", ["base64_n4", "base64_n32"])
def test_fp32_step_vs_fp64_oracle(name):
    g = Golden(name)
    st = g.state()
    r64 = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight, dtype=torch.float64)
    r32 = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight)
    m = build_model(g.cfg, st, "fp32")
    res = train_step(m, g.cfg, g.x, g.target, g.eps, g.ce_weight)
    assert abs(res.loss - r64.loss) <= 1e-5 * abs(r64.loss)
    worst_ours = worst_ref = 0.0
    bad = []
    for n, _ in O.param_specs(g.cfg):
        if n == "decoder.conv2.bias":
            continue
        e_ours = rel_l2(res.grads[n], r64.grads[n])
        e_ref = rel_l2(r32.grads[n], r64.grads[n])
        worst_ours, worst_ref = max(worst_ours, e_ours), max(worst_ref, e_ref)
        if e_ours > max(1e-5, 3 * e_ref):
            bad.append((n, e_ours, e_ref))
    print(f"{name}: worst grad rel-L2 vs fp64: ours {worst_ours:.2e}, torch-fp32 {worst_ref:.2e}")
    assert not bad, bad[:10]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_layer_activations_vs_oracle(prec):
    """Every raw conv output and block output of the forward pass, layer by layer."""
    g = Golden("base64_n4")
    st = g.state()
    mu, lv, enc, recon, ctx = O.forward(st, g.cfg, g.x, g.eps, training=True, keep_activations=True)
    m = build_model(g.cfg, st, prec)
    m.train(True)
    with torch.no_grad():
        m._run_forward(g.x.cuda(), g.eps.cuda(), training=True)
    torch.cuda.synchronize()
    tol = 1e-4 if prec == "fp32" else 3e-2
    bad = []
    for name, ref in ctx.acts.items():
        got = workspace_tensor(m, g.x.shape[0], name)
        e = rel_l2(got, ref)
        if not e <= tol:
            bad.append((name, e))
    assert not bad, bad


@pytest.mark.parametrize("name", ["base64_n4", "base64_n32", "categorical2_n2", "size32_z32_n3"])
def test_bf16_step_within_bf16_floor(name):
    """bf16 mode.  Loss: 1e-2 relative against the fp64 oracle (north_star).  Gradients: a 1e-2 per-tensor bound
    is not attainable by ANY implementation that keeps activations in bf16 on this network -- a bf16-rounded
    forward flips a fraction f ~ 1 % of ReLU masks and each flip changes its gradient entry by O(1), so per-tensor
    gradients sit sqrt(f) ~ 10-40 % (relative L2) from the fp32 ones; the CPU oracle with bf16 rounding inserted
    at the storage points (emulate_bf16) shows the same ('bf16 parity').  The test therefore bounds
    the CUDA path by that intrinsic floor: no tensor worse than 2x the emulated-bf16 distance from fp64 (+0.1),
    same median, and every gradient still points the same way (cosine > 0.85, or within 0.1 of the emulated floor)."""
    g = Golden(name)
    st = g.state()
    ref = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight, dtype=torch.float64)
    emu = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight, dtype=torch.float64, emulate_bf16=True)
    m = build_model(g.cfg, st, "bf16")
    res = train_step(m, g.cfg, g.x, g.target, g.eps, g.ce_weight)
    assert abs(res.loss - ref.loss) <= 1e-2 * abs(ref.loss)
    names = [n for n, _ in O.param_specs(g.cfg) if n != "decoder.conv2.bias"]
    ours = {n: rel_l2(res.grads[n], ref.grads[n]) for n in names}
    floor = {n: rel_l2(emu.grads[n], ref.grads[n]) for n in names}
    cos = {n: torch.nn.functional.cosine_similarity(res.grads[n].double().reshape(1, -1),
                                                    ref.grads[n].reshape(1, -1)).item() for n in names}
    mo, mf = float(np.median(list(ours.values()))), float(np.median(list(floor.values())))
    print(f"{name}: bf16 grad rel-L2 vs fp64: ours median {mo:.3f} max {max(ours.values()):.3f}; "
          f"emulated-bf16 oracle median {mf:.3f} max {max(floor.values()):.3f}; min cosine {min(cos.values()):.3f}")
    # the emulated floor is ONE draw of the same chaos (which gates flip depends on rounding at the 1e-7 level): on the 2-3
    # frame fixtures a single tensor lands up to 0.1 either side of twice the oracle's draw; the rigorous per-kernel bound is
    # tests/test_gpu_bwd_referee.py
    bad = {n: (ours[n], floor[n]) for n in names if ours[n] > 2 * floor[n] + 0.1}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0])[:10]
    assert mo <= 1.5 * mf + 0.01
    cos_floor = min(torch.nn.functional.cosine_similarity(emu.grads[n].reshape(1, -1), ref.grads[n].reshape(1, -1)).item()
                    for n in names)
    assert min(cos.values()) > min(0.85, cos_floor - 0.1), sorted(cos.items(), key=lambda kv: kv[1])[:5]
    assert rel_l2(res.recon, ref.recon) <= 2 * rel_l2(emu.recon, ref.recon) + 1e-3
    assert rel_l2(res.mu, ref.mu) <= 2 * rel_l2(emu.mu, ref.mu) + 1e-3


def test_bench_batch_bf16_against_fp32_mode():
    """The bench configuration (N=256, 64x64): the bf16 product path against the fp32 validation path of the same
    library (itself pinned to the reference fixtures above) on the same weights, frames and noise.  Loss within 1e-2;
    gradients within the bf16-storage floor seen on the small fixtures (cosine > 0.85, relative L2 < 0.6)."""
    cfg = O.VAEConfig(input_image_size=64, z_dimension=64)
    st = O.init_state(cfg, seed=3)
    n = 256
    x = O.normalise(O.synthetic_labels(n, 64))
    eps = torch.randn(n, 64, 1, 1, generator=torch.Generator().manual_seed(11))
    r32 = train_step(build_model(cfg, st, "fp32"), cfg, x, x, eps)
    r16 = train_step(build_model(cfg, st, "bf16"), cfg, x, x, eps)
    assert abs(r16.loss - r32.loss) <= 1e-2 * abs(r32.loss)
    names = [k for k, _ in O.param_specs(cfg) if k != "decoder.conv2.bias"]
    cos = {k: torch.nn.functional.cosine_similarity(r16.grads[k].double().reshape(1, -1),
                                                    r32.grads[k].double().reshape(1, -1)).item() for k in names}
    rel = {k: rel_l2(r16.grads[k], r32.grads[k]) for k in names}
    print(f"N=256: bf16 vs fp32 mode: min cosine {min(cos.values()):.3f}, max rel-L2 {max(rel.values()):.3f}, "
          f"median rel-L2 {float(np.median(list(rel.values()))):.3f}")
    assert min(cos.values()) > 0.85, sorted(cos.items(), key=lambda kv: kv[1])[:5]
    assert max(rel.values()) < 0.6, sorted(rel.items(), key=lambda kv: -kv[1])[:5]
    assert rel_l2(r16.recon, r32.recon) < 3e-2 and rel_l2(r16.mu, r32.mu) < 3e-2


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_widened_vae_against_oracle(prec):
    """BASELINE configs[3]: the widened VAE (2x channels, latent dim 256), here at N=4.  fp32 mode against the fp64
    oracle within 3x the distance of a torch-fp32 evaluation from it; bf16 mode: loss within 1e-2 and gradients along the oracle's."""
    cfg = O.VAEConfig(input_image_size=64, z_dimension=256, width=2)
    st = O.init_state(cfg, seed=5)
    n = 4
    x = O.normalise(O.synthetic_labels(n, 64))
    eps = torch.randn(n, 256, 1, 1, generator=torch.Generator().manual_seed(12))
    ref = O.train_step(st, cfg, x, x, eps, dtype=torch.float64)
    res = train_step(build_model(cfg, st, prec), cfg, x, x, eps)
    names = [k for k, _ in O.param_specs(cfg) if k != "decoder.conv2.bias"]
    if prec == "fp32":
        # same two-sided bar as test_fp32_step_vs_fp64_oracle: no worse than 3x torch-fp32's own distance from fp64
        r32 = O.train_step(st, cfg, x, x, eps)
        assert abs(res.loss - ref.loss) <= 1e-5 * abs(ref.loss)
        bad = [(k, rel_l2(res.grads[k], ref.grads[k]), rel_l2(r32.grads[k], ref.grads[k])) for k in names
               if rel_l2(res.grads[k], ref.grads[k]) > max(1e-5, 3 * rel_l2(r32.grads[k], ref.grads[k]))]
        assert not bad, bad[:10]
    else:
        assert abs(res.loss - ref.loss) <= 1e-2 * abs(ref.loss)
        cos = {k: torch.nn.functional.cosine_similarity(res.grads[k].double().reshape(1, -1),
                                                        ref.grads[k].double().reshape(1, -1)).item() for k in names}
        assert min(cos.values()) > 0.8, sorted(cos.items(), key=lambda kv: kv[1])[:5]


def test_eval_mode_and_decoder_only():
    g = Golden("base64_n4")
    st = g.state()
    # make the running statistics non-trivial
    gen = torch.Generator().manual_seed(5)
    for k in st:
        if k.endswith("running_mean"):
            st[k] = 0.1 * torch.randn(st[k].shape, generator=gen)
        elif k.endswith("running_var"):
            st[k] = 0.5 + torch.rand(st[k].shape, generator=gen)
    mu, lv, enc, recon, _ = O.forward(st, g.cfg, g.x, g.eps, training=False)
    m = build_model(g.cfg, st, "fp32")
    m.eval()
    with torch.no_grad():
        mu2, lv2, enc2, recon2 = m(g.x.cuda(), eps=g.eps.cuda())
        rec3 = m.get_reconstruction(enc.cuda())
    assert rel_l2(mu2.cpu(), mu) < 1e-4 and rel_l2(recon2.cpu(), recon) < 1e-4
    assert rel_l2(rec3.cpu(), recon) < 1e-4
    # eval must not touch the buffers
    sd = m.state_dict()
    assert torch.equal(sd["encoder.bn1.running_mean"].cpu(), st["encoder.bn1.running_mean"])
    assert int(sd["encoder.bn1.num_batches_tracked"]) == 0


def test_philox_eps_is_reproducible_and_normal():
    import mmvae_b200 as M
    from ctypes import c_void_p
    n = 1 << 16
    a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
    s = c_void_p(torch.cuda.current_stream().cuda_stream)
    M._lib.check(M._lib.lib.mmvae_philox_normal(1234, 0, None, 0, n, c_void_p(a.data_ptr()), s))
    M._lib.check(M._lib.lib.mmvae_philox_normal(1234, 0, None, 0, n, c_void_p(b.data_ptr()), s))
    assert torch.equal(a, b)
    assert abs(a.mean().item()) < 0.02 and abs(a.std().item() - 1.0) < 0.02
    # the module's own draw is exposed so the same noise can be fed to the reference
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "fp32")
    m.train(True)
    mu, lv, enc, rec = m(g.x.cuda())
    eps = m.last_eps
    assert eps.shape == (4, 64, 1, 1)
    assert rel_l2((mu + eps * torch.exp(0.5 * lv)).detach().cpu(), enc.detach().cpu()) < 1e-6


def test_kl_weight_and_kl_divergence():
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "fp32")
    mu = torch.randn(4, 64, 1, 1, device="cuda", requires_grad=True)
    lv = (0.3 * torch.randn(4, 64, 1, 1, device="cuda")).requires_grad_(True)
    kl = m.kl_divergence(mu, lv)
    want = O.kl_sum(mu.detach().cpu().double(), lv.detach().cpu().double())
    assert abs(kl.item() - want.item()) <= 1e-5 * abs(want.item())
    kl.backward()
    assert rel_l2(mu.grad.cpu(), mu.detach().cpu()) < 1e-5                       # dKL/dmu = mu
    assert rel_l2(lv.grad.cpu(), 0.5 * (torch.exp(lv.detach().cpu()) - 1)) < 1e-5


@pytest.mark.parametrize("kw,n", [(dict(input_image_size=64, z_dimension=64), 256),                 # BASELINE configs[1]
                                  (dict(input_image_size=64, z_dimension=256, width=2), 128)])     # configs[3], per-GPU share
def test_fp32_mode_at_bench_sizes_vs_fp64_oracle(kw, n):
    """The BENCH batch sizes pinned to the oracle: fp32 validation mode of the library against an fp64 evaluation of
    the reference's formulas on the same frames, weights and noise.  Loss 1e-5; every gradient no worse than 3x
    torch-fp32's own distance from fp64, with the fixtures' 5e-4 (the reference-fp32 noise floor the golden tests use) as
    the floor: at these sizes single tensors of either fp32 evaluation land anywhere between 1e-5 and 5e-3 from fp64."""
    cfg = O.VAEConfig(**kw)
    st = O.init_state(cfg, seed=3)
    x = O.normalise(O.synthetic_labels(n, 64))
    eps = torch.randn(n, cfg.z_dimension, 1, 1, generator=torch.Generator().manual_seed(11))
    r64 = O.train_step(st, cfg, x, x, eps, dtype=torch.float64)
    r32 = O.train_step(st, cfg, x, x, eps)
    res = train_step(build_model(cfg, st, "fp32"), cfg, x, x, eps)
    assert abs(res.loss - r64.loss) <= 1e-5 * abs(r64.loss)
    assert rel_l2(res.recon, r64.recon) <= 1e-4 and rel_l2(res.mu, r64.mu) <= 1e-4
    bad, worst, worst_ref = [], 0.0, 0.0
    for k, _ in O.param_specs(cfg):
        if k == "decoder.conv2.bias":
            continue
        e, e_ref = rel_l2(res.grads[k], r64.grads[k]), rel_l2(r32.grads[k], r64.grads[k])
        worst, worst_ref = max(worst, e), max(worst_ref, e_ref)
        if e > max(5e-4, 3 * e_ref):
            bad.append((k, e, e_ref))
    print(f"width {cfg.width} N={n}: worst grad rel-L2 vs fp64: ours {worst:.2e}, torch-fp32 {worst_ref:.2e}")
    assert not bad, bad[:10]
    for prefix, _ in O.bn_names(cfg):
        assert rel_l2(res.new_buffers[prefix + ".running_var"], r64.new_buffers[prefix + ".running_var"]) <= 1e-4
