"""The layer-local backward referee (oracle.local_backward, the judge of the bf16 CUDA backward in tests/test_gpu_bwd_referee.py
and smoke()) against autograd on the CPU: fed with the activations and activation gradients of ONE fp64 evaluation it must
reproduce that evaluation's parameter gradients and gradient tensors to rounding -- and, when the gradient of the raw stem
output is withheld (the CUDA path forms that dY inside the stem weight-gradient kernel's loader and never stores it), judge the
stem weight gradient on its own expected dY rounded to bf16."""
import torch

from oracle import vae_oracle as O


def _evaluation(n=2):
    cfg = O.VAEConfig(input_image_size=64, z_dimension=64)
    st = O.init_state(cfg, seed=3)
    names = [k for k, _ in O.param_specs(cfg)]
    # conv weights as the tensor cores see them, so that the referee's bf16 operands are this evaluation's weights
    st = {k: (v.to(torch.bfloat16).to(v.dtype) if (k in names and v.dim() == 4 and k not in O.FP32_OPERANDS) else v) for k, v in st.items()}
    x = O.normalise(O.synthetic_labels(n, 64))
    eps = torch.randn(n, 64, 1, 1, generator=torch.Generator().manual_seed(5))
    work = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in st.items()}
    for k in names:
        work[k].requires_grad_(True)
    mu, logvar, enc, recon, ctx = O.forward(work, cfg, x.double(), eps.double(), training=True, keep_activations=True)
    total, _, _ = O.loss(cfg, x.double(), mu, logvar, recon)
    acts = dict(ctx.acts)
    keys = list(acts)
    got = torch.autograd.grad(total, [work[k] for k in names] + [acts[k] for k in keys], allow_unused=True)
    pg = {k: g for k, g in zip(names, got[:len(names)])}
    ag = {k: (g if g is not None else torch.zeros_like(acts[k])) for k, g in zip(keys, got[len(names):])}
    fwd = {k: v.detach() for k, v in acts.items()}
    d_recon = cfg.nll * (recon.detach() - x.double()) / (cfg.sigma_decoder ** 2) / n
    d_mu = (cfg.kl / n) * mu.detach()
    d_lv = (cfg.kl / n) * 0.5 * (torch.exp(logvar.detach()) - 1)
    return cfg, st, x, eps, fwd, ag, pg, d_recon, d_mu, d_lv


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def test_local_backward_reproduces_autograd():
    cfg, st, x, eps, fwd, ag, pg, d_recon, d_mu, d_lv = _evaluation()
    P, A = O.local_backward(st, cfg, x, eps, fwd, dict(ag), d_recon, d_mu, d_lv)
    for k, want in P.items():
        if k == "decoder.conv2.bias":
            continue
        assert _rel(want, pg[k]) <= 1e-9, (k, _rel(want, pg[k]))
    for k, (want, gate) in A.items():
        got = ag[k]
        if gate is not None:
            got, want = got * gate, want * gate
        assert _rel(want, got) <= 1e-9, (k, _rel(want, got))


def test_local_backward_without_the_stored_stem_gradient():
    cfg, st, x, eps, fwd, ag, pg, d_recon, d_mu, d_lv = _evaluation()
    grd = {k: v for k, v in ag.items() if k != "encoder.conv1"}
    P, A = O.local_backward(st, cfg, x, eps, fwd, grd, d_recon, d_mu, d_lv)
    # judged on the referee's own dY, rounded to bf16 (2^-9 relative per element, independent): well inside 1e-2
    e = _rel(P["encoder.conv1.weight"], pg["encoder.conv1.weight"])
    assert 0 < e <= 5e-3, e
    # everything else is untouched by the missing tensor
    for k in ("encoder.bn1.weight", "encoder.bn1.bias", "encoder.layer1.0.conv1.weight"):
        assert _rel(P[k], pg[k]) <= 1e-9, k
    assert _rel(A["encoder.conv1"][0], ag["encoder.conv1"]) <= 1e-9
