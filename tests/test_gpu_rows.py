"""GPU: the rows around the hot path and the module's edge cases, through the
public API -> C ABI, against the oracle / the fixtures made from the live reference:

  * fused Adam (`mmvae_b200.FusedAdam`, main.py:468,399) vs torch.optim.Adam on identical gradients, vs the fixture of
    the reference's optimizer, and the whole loop body main.py:389-399 over three iterations vs the reference;
  * the MMD diagnostic (model.py:367-383,394-396) vs the reference's value on the same true_samples;
  * device-side input normalisation (main.py:381-388) bit-exact vs the oracle's `normalise`;
  * `VAE.loss(kl_weight=...)` on the model.py path, `defer_metrics`, the KL-only backward of a cropped model,
    phased backward == one-call backward, the double-backward guard, target validation;
  * GraphedTrainStep: no trace of the warm-up in the module, gradients re-attached after zero_grad, KL annealing
    through the device scalar, the captured optimizer step.
"""
import os
import types

import numpy as np
import pytest
import torch

import mmvae_b200 as M
from golden_util import GOLDEN_DIR, Golden
from mmvae_b200 import _lib
from mmvae_b200 import data as D
from oracle import vae_oracle as O
from ours_util import build_model, rel_l2, train_step
from test_oracle_aux import check_final_state

pytestmark = pytest.mark.gpu
NS = types.SimpleNamespace


# ------------------------------------------------------------------ fused Adam
def test_fused_adam_matches_torch_optim_adam():
    """Ten steps on the real gradient arena of the model (fp32 mode), same gradients fed to torch.optim.Adam."""
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "fp32")
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters()]
    ref_opt = torch.optim.Adam(ref_params, lr=1e-3)
    opt = M.FusedAdam(m, lr=1e-3)
    x = g.x.cuda()
    gen = torch.Generator().manual_seed(21)
    for it in range(10):
        eps = torch.randn(g.eps.shape, generator=gen).cuda()
        mu, lv, enc, rec = m(x, eps=eps)
        loss, *_ = m.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None))
        opt.zero_grad()
        loss.backward()
        for rp, p in zip(ref_params, m.parameters()):
            rp.grad = p.grad.detach().clone()
        opt.step()
        ref_opt.step()
        worst = max((p.detach() - rp.detach()).abs().max().item() for rp, p in zip(ref_params, m.parameters()))
        assert worst <= 2e-6, (it, worst)                      # lr = 1e-3: a fraction of a thousandth of one step
    torch.cuda.synchronize()
    st = ref_opt.state[ref_params[0]]
    assert torch.allclose(opt.exp_avg[:ref_params[0].numel()].view_as(st["exp_avg"]), st["exp_avg"], rtol=2e-4, atol=1e-9)
    assert opt.step_count == 10 and int(opt._step_dev) == 10


def test_fused_adam_fixture_and_weight_decay():
    """The optimizer alone against the trajectory recorded from torch.optim.Adam (tests/golden/aux_adam.npz)."""
    z = np.load(os.path.join(GOLDEN_DIR, "aux_adam.npz"))

    class Flat(torch.nn.Module):
        def __init__(self, p0):
            super().__init__()
            self.p = torch.nn.Parameter(p0.clone())
            self.flat_parameters = self.p.data
            self.last_flat_grad = None

    for tag, wd in (("plain", 0.0), ("wd", 0.01)):
        mod = Flat(torch.from_numpy(z["opt/p0"]).cuda())
        opt = M.FusedAdam(mod, lr=float(z["lr"]), weight_decay=wd)
        for it in range(5):
            opt.step(torch.from_numpy(z["opt/grads"][it]).cuda())
            want = torch.from_numpy(z[f"opt/{tag}/traj"][it])
            assert (mod.p.detach().cpu() - want).abs().max().item() <= 2e-6, (tag, it)
        assert torch.allclose(opt.exp_avg_sq.cpu(), torch.from_numpy(z[f"opt/{tag}/exp_avg_sq"]), rtol=2e-4, atol=1e-9)
    # grad_scale: a summed (not averaged) exchange over 4 ranks
    mod = Flat(torch.from_numpy(z["opt/p0"]).cuda())
    opt = M.FusedAdam(mod, lr=float(z["lr"]), grad_scale=0.25)
    opt.step(4.0 * torch.from_numpy(z["opt/grads"][0]).cuda())
    assert (mod.p.detach().cpu() - torch.from_numpy(z["opt/plain/traj"][0])).abs().max().item() <= 2e-6


def test_loop_body_with_adam_matches_reference_trajectory():
    """main.py:389-399 three times (forward, loss, zero_grad, backward, Adam) in fp32 mode against the reference's own
    losses and final state (tests/golden/aux_adam.npz), eager and as ONE captured graph per iteration."""
    z = np.load(os.path.join(GOLDEN_DIR, "aux_adam.npz"))
    g = Golden("base64_n4")
    lr = float(z["lr"])
    # eager
    m = build_model(g.cfg, g.state(), "fp32")
    opt = M.FusedAdam(m, lr=lr)
    x = g.x.cuda()
    losses = []
    for e in z["eps"]:
        mu, lv, enc, rec = m(x, eps=torch.from_numpy(e).cuda())
        loss, *_ = m.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    ref = z["losses"]
    assert abs(losses[0] - ref[0]) <= 1e-5 * abs(ref[0])
    for a, b in zip(losses[1:], ref[1:]):
        assert abs(a - b) <= 2e-3 * abs(b), (losses, ref)
    check_final_state(m.state_dict(), g.state(), z, lr, 3)
    # graph-captured step with the optimizer inside: same trajectory as a host loop on the noise the graph drew
    m1 = build_model(g.cfg, g.state(), "fp32")
    opt1 = M.FusedAdam(m1, lr=lr)
    step = M.GraphedTrainStep(m1, x.shape[0], warmup=2, optimizer=opt1)
    m2 = build_model(g.cfg, g.state(), "fp32")
    opt2 = M.FusedAdam(m2, lr=lr)
    for it in range(3):
        l1, _, _ = step(x)
        torch.cuda.synchronize()
        mu, lv, enc, rec = m2(x, eps=m1.last_eps.clone())
        l2, *_ = m2.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None))
        opt2.zero_grad()
        l2.backward()
        opt2.step()
        assert abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l2)), (it, float(l1), float(l2))
    assert opt1.step_count == 3 and int(opt1._step_dev) == 3
    # Adam moves an entry by ~lr per step whatever the size of its gradient: the fp32 atomics of the weight-gradient
    # reductions (order differs between a replay and host launches) flip the sign of a few near-zero entries, each worth
    # 2e-3 after three steps -- 1e-3 relative L2 allows ~0.1 % of the 2.1 M entries to differ by a whole step
    assert rel_l2(m1.flat_parameters, m2.flat_parameters) <= 1e-3
    for k, v in m2.state_dict().items():
        if k.endswith("running_var"):
            assert rel_l2(m1.state_dict()[k], v) <= 1e-3


# ------------------------------------------------------------------ MMD diagnostic
def test_mmd_diagnostic_matches_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "aux_mmd.npz"))
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "fp32")
    got2 = m.compute_mmd(torch.from_numpy(z["x2"]).cuda(), torch.from_numpy(z["y2"]).cuda()).item() * z["x2"].shape[0]
    assert abs(got2 - float(z["mmd2"])) <= 1e-4 * abs(float(z["mmd2"]))      # a difference of three sums ~230x its size
    # through loss(): the reference's 4th return value on the reference's own true_samples draw
    x = g.x.cuda()
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    ts = torch.from_numpy(z["true_samples"]).cuda()
    loss, pxz, kl, mmd = m.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None), true_samples=ts)
    assert abs(mmd - float(z["mmd_over_n"])) <= 1e-3 * abs(float(z["mmd_over_n"]))
    assert abs(pxz - float(g.z["pxz"])) <= 5e-4 * abs(float(g.z["pxz"]))
    # default: a Philox draw the caller can read back and feed to the reference's compute_mmd
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    _, _, _, mmd2 = m.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None))
    want = O.compute_mmd(m.last_true_samples.cpu().double(), enc.detach().cpu().double().view(4, -1)).item() / 4
    assert abs(mmd2 - want) <= 1e-4 * abs(want)
    assert abs(m.last_true_samples.mean().item()) < 0.3 and abs(m.last_true_samples.std().item() - 1) < 0.2
    # encoding None (model.py:394): no MMD
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    assert m.loss(x, mu, lv, None, rec, x.device, NS(data_ratio_of_labels=None))[3] == 0.0


# ------------------------------------------------------------------ input side
@pytest.mark.parametrize("n,size", [(7, 64), (256, 64), (3, 28)])
def test_prepare_input_bit_exact_vs_oracle(n, size):
    labels = O.synthetic_labels(n, size)
    want = O.normalise(labels)                                   # main.py:383-388 on the CPU
    x, tgt = D.prepare_input(labels.cuda(), want_target=True)
    assert torch.equal(x.cpu(), want)                            # bit-exact: same fp32 subtract and divide
    assert torch.equal(tgt.cpu(), labels.long())                 # main.py:382
    # a k-means label map with more than two labels (k = 4), non-default statistics
    lab4 = torch.randint(0, 4, (5, size, size), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    x4 = D.prepare_input(lab4.cuda(), 1.37, 0.81)
    assert torch.equal(x4.cpu(), (lab4.float().view(5, 1, size, size) - 1.37) / 0.81)


# ------------------------------------------------------------------ loss options
@pytest.mark.parametrize("name", ["base64_n4", "categorical2_n2"])
def test_loss_kl_weight_model_py_path(name):
    g = Golden(name)
    st = g.state()
    ref = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight, kl_weight=0.35, dtype=torch.float64)
    res = train_step(build_model(g.cfg, st, "fp32"), g.cfg, g.x, g.target, g.eps, g.ce_weight, kl_weight=0.35)
    assert abs(res.loss - ref.loss) <= 1e-5 * abs(ref.loss)
    assert abs(res.kl - ref.kl) <= 1e-5 * abs(ref.kl) and abs(res.pxz - ref.pxz) <= 1e-5 * abs(ref.pxz)
    r32 = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight, kl_weight=0.35)
    for k, _ in O.param_specs(g.cfg):
        if k != "decoder.conv2.bias":
            assert rel_l2(res.grads[k], ref.grads[k]) <= max(1e-5, 3 * rel_l2(r32.grads[k], ref.grads[k])), k


def test_defer_metrics_equals_host_floats():
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "fp32")
    x = g.x.cuda()
    ts = torch.randn(4, 64, generator=torch.Generator().manual_seed(1)).cuda()
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    l1, pxz1, kl1, mmd1 = m.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None), true_samples=ts)
    assert all(isinstance(v, float) for v in (pxz1, kl1, mmd1))
    m.defer_metrics = True
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    l2, pxz2, kl2, mmd2 = m.loss(x, mu, lv, enc, rec, x.device, NS(data_ratio_of_labels=None), true_samples=ts)
    assert all(torch.is_tensor(v) and v.is_cuda and v.dim() == 0 for v in (pxz2, kl2, mmd2))
    assert (float(l1), pxz1, kl1, mmd1) == (float(l2), float(pxz2), float(kl2), float(mmd2))


def test_kl_only_backward_on_cropped_model():
    """A loss without a reconstruction term on a configuration whose decoder output is cropped (S=28 -> D=32):
    d_recon goes through as NULL and the decoder sweep is skipped."""
    g = Golden("crop28_n2")
    st = g.state()
    m = build_model(g.cfg, st, "fp32")
    x = g.x.cuda()
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    m.kl_divergence(mu, lv).backward()
    torch.cuda.synchronize()
    names = [n for n, _ in O.param_specs(g.cfg)]
    work = {k: v.clone().double() for k, v in st.items()}
    for n in names:
        work[n].requires_grad_(True)
    omu, olv, _, _, _ = O.forward(work, g.cfg, g.x.double(), g.eps.double())
    grads = torch.autograd.grad(O.kl_sum(omu, olv), [work[n] for n in names], allow_unused=True)
    for n, gr in zip(names, grads):
        got = dict(m.named_parameters())[n].grad
        if gr is None or n.startswith("decoder."):
            assert got.abs().max().item() == 0.0, n
        else:
            assert rel_l2(got.cpu(), gr) <= 2e-3, (n, rel_l2(got.cpu(), gr))


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_phased_backward_equals_one_call(prec):
    """The three-phase backward the data-parallel path uses (DECODER, ENC_DEEP, ENC_SHALLOW with a hand-over after
    each) against MMVAE_BWD_ALL on one GPU."""
    class NoSync:                                   # GradSync's interface, world of one
        def __init__(self):
            self.phases = []

        def phase_done(self, model, desc, grads, phase):
            self.phases.append(phase)

        def finish(self):
            pass

    g = Golden("base64_n32")
    a = train_step(build_model(g.cfg, g.state(), prec), g.cfg, g.x, g.x, g.eps)
    m = build_model(g.cfg, g.state(), prec)
    m._grad_sync = NoSync()
    b = train_step(m, g.cfg, g.x, g.x, g.eps)
    assert m._grad_sync.phases == [_lib.BWD_DECODER, _lib.BWD_ENC_DEEP, _lib.BWD_ENC_SHALLOW]
    tol = 1e-5 if prec == "fp32" else 2e-3          # fp32 atomics of the weight-gradient reductions reorder sums
    bad = [(k, rel_l2(b.grads[k], a.grads[k])) for k in a.grads
           if k != "decoder.conv2.bias" and rel_l2(b.grads[k], a.grads[k]) > tol]
    assert not bad, bad[:5]


def test_guards():
    g = Golden("categorical2_n2")
    m = build_model(g.cfg, g.state(), "fp32")
    x = g.x.cuda()
    mu, lv, enc, rec = m(x, eps=g.eps.cuda())
    args = NS(data_ratio_of_labels=None)
    with pytest.raises(ValueError):                 # wrong number of class indices
        m.loss(g.target[:, :32].cuda(), mu, lv, enc, rec, x.device, args)
    with pytest.raises(ValueError):
        m.loss(g.target[:1].cuda(), mu, lv, enc, rec, x.device, args)
    bad = g.target.clone()
    bad[0, 0, 0] = 2                                # class outside [0, C): NaN loss, never an out-of-bounds read
    loss, *_ = m.loss(bad.cuda(), mu, lv, enc, rec, x.device, args)
    assert torch.isnan(loss).item()
    loss, *_ = m.loss(g.target.cuda(), mu, lv, enc, rec, x.device, args)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="twice"):
        loss.backward()


# ------------------------------------------------------------------ GraphedTrainStep side effects
def test_graphed_step_leaves_no_trace_and_reattaches_grads():
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "bf16")
    before = {k: v.clone() for k, v in m.state_dict().items()}
    defer = m.defer_metrics
    step = M.GraphedTrainStep(m, 4, warmup=3)
    torch.cuda.synchronize()
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), f"the warm-up left a trace in {k}"
    assert m.defer_metrics == defer
    x = g.x.cuda()
    step(x)
    torch.cuda.synchronize()
    g1 = m.flat_parameters.new_tensor([p.grad.abs().sum().item() for p in m.parameters()])
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.zero_grad()                                 # set_to_none: detaches the graph's static gradient tensors
    assert all(p.grad is None for p in m.parameters())
    second = M.GraphedTrainStep(m, 4, warmup=1)     # and a second graph binds p.grad to ITS buffers
    step(x)
    torch.cuda.synchronize()
    assert all(p.grad is not None and p.grad.data_ptr() == sg.data_ptr() for p, sg in zip(m.parameters(), step.grads))
    g2 = m.flat_parameters.new_tensor([p.grad.abs().sum().item() for p in m.parameters()])
    assert (g2 > 0).sum() >= (g1 > 0).sum() - 1 and torch.isfinite(g2).all()
    w0 = m.flat_parameters.clone()
    opt.step()                                      # a stock optimizer sees the replay's gradients
    assert not torch.equal(w0, m.flat_parameters)
    del second


def test_graphed_step_kl_annealing():
    g = Golden("base64_n4")
    m = build_model(g.cfg, g.state(), "fp32")
    x = g.x.cuda()
    step = M.GraphedTrainStep(m, 4, warmup=1, kl_weight=1.0)
    vals = {}
    for w in (1.0, 0.25, 0.0):
        # same weights and BatchNorm state every time (no optimizer); the noise differs per replay, so compare through
        # the components the graph returns: loss == pxz + w * kl
        step.set_kl_weight(w)
        loss, pxz, kl = step(x)
        torch.cuda.synchronize()
        vals[w] = (float(loss), float(pxz), float(kl))
        assert abs(float(loss) - (float(pxz) + w * float(kl))) <= 1e-5 * abs(float(loss)), (w, vals[w])
    assert vals[0.0][2] > 0


def test_backward_chain_is_reproducible():
    """Two bf16 steps on identical inputs: every stored gradient tensor of the backward chain is BIT-identical (fixed
    reduction orders; fp64 accumulation across CTAs), so the run-to-run difference of the parameter gradients is the fp32
    atomics of the weight-gradient reductions alone.  A 1e-7 wobble in one BatchNorm-backward sum would be amplified to
    ~1e-2 at the stem by the bf16 rounding of the ~25 gradient tensors downstream (it was, with shared-memory float atomics)."""
    from ours_util import workspace_tensor
    g = Golden("base64_n32")
    runs = []
    for _ in range(2):
        m = build_model(g.cfg, g.state(), "bf16")
        res = train_step(m, g.cfg, g.x, g.x, g.eps)
        chain = {k: workspace_tensor(m, g.x.shape[0], k + ".grad") for k in
                 ("decoder.uplayer3.0.conv2", "decoder.input", "encoder.layer4.0", "encoder.layer2.0.conv1", "encoder.relu")}   # (the stem's own dY is formed in the weight-gradient loader, never stored)
        runs.append((res, chain))
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), f"{k}.grad differs between two identical steps"
    for k, ga in runs[0][0].grads.items():
        if k != "decoder.conv2.bias":
            assert rel_l2(runs[1][0].grads[k], ga) <= 1e-5, k
