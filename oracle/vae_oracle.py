"""CPU oracle for the VAE training step (forward + loss + backward).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mmvae_b200/`` may import this file.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the thing timed on
the host cores -- never as the product path.

What it restates (reference = /root/reference, praateekmahajan/moving-mnist-vae):

  * ``VAE_Encoder``  model.py:88-150   (stem conv k5 s2 p2 + BN + ReLU, four
    ``BasicBlock`` model.py:23-55 with a 1x1/BN shortcut, avg-pool, 1x1 heads)
  * ``rsample``      model.py:148-150  (mu + eps * exp(0.5 * logvar))
  * ``VAE_Decoder``  model.py:153-209  (convT k2 stem + BN + ReLU, four or five
    ``DeconvBottleneck`` model.py:57-85, conv 3x3 + bias + BN tail)
  * ``VAE.forward``  model.py:316-342  (crop by ``adjust`` model.py:307-310)
  * ``VAE.loss``     model.py:385-406  (Gaussian NLL model.py:403 or weighted
    cross-entropy model.py:399-401, KL model.py:364-365, /N model.py:405)
  * BatchNorm2d training semantics relied on by all of the above (SURVEY.md
    Appendix A): biased batch variance for normalisation, unbiased variance
    and momentum 0.1 for the running buffers.

The arithmetic itself lives in PyTorch (third-party for the reference, not
vendored, not pinned by it; torch 2.11.0 is what this image has).  The oracle is
written functionally over a ``state_dict`` (same key names as the reference) so
that it has no ``nn.Module`` in common with the reference source: convolutions
go through ``torch.nn.functional`` on the CPU, BatchNorm / loss / KL are spelled
out as formulas, gradients come from autograd over those formulas.

Parity pin: the reference stores NO golden vectors, so
this oracle is pinned against outputs of the reference itself, produced in the
build container by ``tests/golden/make_golden.py`` (imports
/root/reference/model.py unmodified) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them.  The fixtures are self-generated,
and DESIGN.md says so.

Extensions the reference cannot express: ``width`` multiplies
the hard-coded channel literals of model.py:92-101 / model.py:157-170 (config
C4); ``kl_weight`` may be passed per call (the "annealed" configuration).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, constructed at model.py:30 etc.
BN_MOMENTUM = 0.1


@dataclass(frozen=True)
class VAEConfig:
    """Constructor arguments of the reference ``VAE`` (model.py:259-262) that the
    hot path depends on, plus the builder-defined ``width`` multiplier."""
    in_channels: int = 1
    decoder_out_channels: int = 1
    z_dimension: int = 64
    input_image_size: int = 64
    nll: float = 1.0
    kl: float = 1.0
    sigma_decoder: float = 0.1
    require_rsample: bool = True
    width: int = 1

    @property
    def categorical(self) -> bool:
        # model.py:399: CE branch when decoder_out_channels > in_channels
        return self.decoder_out_channels > self.in_channels

    @property
    def enc_planes(self) -> Tuple[int, ...]:
        return tuple(c * self.width for c in (32, 64, 128, 256))   # model.py:97-100

    @property
    def stem_planes(self) -> int:
        return 32 * self.width                                      # model.py:92-94

    @property
    def dec_stem_planes(self) -> int:
        return 128 * self.width                                     # model.py:157

    @property
    def dec_planes(self) -> Tuple[int, ...]:
        p = [128, 64, 32, 16]                                       # model.py:164-167
        if self.input_image_size > 32:
            p.append(16)                                            # model.py:169-170
        return tuple(c * self.width for c in p)

    @property
    def adjust(self) -> int:
        # model.py:307-310
        if self.input_image_size > 32:
            return (64 - self.input_image_size) // 2
        return (32 - self.input_image_size) // 2


# ----------------------------------------------------------------------------
# parameter / buffer inventory (same names, shapes and order as the reference's
# named_parameters(); checked against the live reference by the golden script)
# ----------------------------------------------------------------------------

def param_specs(cfg: VAEConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(prefix: str, c: int):
        out.append((prefix + ".weight", (c,)))
        out.append((prefix + ".bias", (c,)))

    s = cfg.stem_planes
    out.append(("encoder.conv1.weight", (s, cfg.in_channels, 5, 5)))
    bn("encoder.bn1", s)
    inpl = s
    for i, pl in enumerate(cfg.enc_planes, start=1):
        p = f"encoder.layer{i}.0"
        out.append((p + ".conv1.weight", (pl, inpl, 3, 3)))
        bn(p + ".bn1", pl)
        out.append((p + ".conv2.weight", (pl, pl, 3, 3)))
        bn(p + ".bn2", pl)
        out.append((p + ".downsample.0.weight", (pl, inpl, 1, 1)))
        bn(p + ".downsample.1", pl)
        inpl = pl
    out.append(("encoder.conv_mu.weight", (cfg.z_dimension, inpl, 1, 1)))
    if cfg.require_rsample:
        out.append(("encoder.conv_logvar.weight", (cfg.z_dimension, inpl, 1, 1)))
    d = cfg.dec_stem_planes
    out.append(("decoder.conv1.weight", (cfg.z_dimension, d, 2, 2)))
    bn("decoder.bn1", d)
    inpl = d
    for i, pl in enumerate(cfg.dec_planes, start=1):
        p = f"decoder.uplayer{i}.0"
        out.append((p + ".conv1.weight", (pl, inpl, 1, 1)))
        bn(p + ".bn1", pl)
        out.append((p + ".conv2.weight", (pl, pl, 4, 4)))          # convT [Cin, Cout, 4, 4]
        bn(p + ".bn2", pl)
        out.append((p + ".upsample.0.weight", (inpl, pl, 4, 4)))   # convT [Cin, Cout, 4, 4]
        bn(p + ".upsample.1", pl)
        inpl = pl
    out.append(("decoder.conv2.weight", (cfg.decoder_out_channels, inpl, 3, 3)))
    out.append(("decoder.conv2.bias", (cfg.decoder_out_channels,)))
    bn("decoder.bn2", cfg.decoder_out_channels)
    return out


def bn_names(cfg: VAEConfig) -> List[Tuple[str, int]]:
    """(prefix, channels) of every BatchNorm2d in state_dict order."""
    res = []
    specs = param_specs(cfg)
    for i, (name, shape) in enumerate(specs):
        if name.endswith(".weight") and len(shape) == 1:
            res.append((name[: -len(".weight")], shape[0]))
    return res


def init_state(cfg: VAEConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Parameters drawn from the reference's default-init *distributions*
    (conv/convT weights U(+-1/sqrt(fan_in)) with
    fan_in = weight.size(1) * kH * kW; BN gamma=1, beta=0; tail bias
    U(+-1/sqrt(fan_in))) plus fresh BN buffers.  Not the reference's RNG stream --
    golden fixtures carry the reference's own draws."""
    g = torch.Generator().manual_seed(seed)
    st: Dict[str, torch.Tensor] = {}
    for name, shape in param_specs(cfg):
        if len(shape) == 4:
            bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            st[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        elif name == "decoder.conv2.bias":
            w = st["decoder.conv2.weight"]
            bound = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
            st[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        elif name.endswith(".weight"):
            st[name] = torch.ones(shape, dtype=dtype)
        else:
            st[name] = torch.zeros(shape, dtype=dtype)
    for prefix, c in bn_names(cfg):
        st[prefix + ".running_mean"] = torch.zeros(c, dtype=dtype)
        st[prefix + ".running_var"] = torch.ones(c, dtype=dtype)
        st[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return st


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------

class _RoundBF16(torch.autograd.Function):
    """Storage rounding of the bf16 mode: the value is rounded to bf16 on the way forward and the
    gradient on the way back (both tensors live in bf16 in the CUDA path)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundGradBF16(torch.autograd.Function):
    """Identity on the way forward, bf16 rounding of the gradient on the way back (gradient storage only)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundBF16Operand(torch.autograd.Function):
    """A tensor-core operand that is kept in fp32 in HBM (a weight): bf16 on the way into the product,
    gradient left in fp32 (the CUDA path accumulates dW in fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


# the 1-channel ends of the network run as fp32 SIMT kernels on fp32 weights in the CUDA bf16 mode
FP32_OPERANDS = ("encoder.conv1.weight", "decoder.conv2.weight")


class _Ctx:
    """Carries the state dict, the training flag and the BN side effects."""

    def __init__(self, st, training: bool, keep: bool, emulate_bf16: bool = False, relu_masks=None, forced_acts=None,
                 emulate_bf16_grads: bool = False):
        self.st = st
        self.training = training
        self.keep = keep
        self.emulate_bf16 = emulate_bf16
        self.relu_masks = relu_masks
        self.forced_acts = forced_acts
        self.emulate_bf16_grads = emulate_bf16_grads      # with forced_acts: round the gradient at every store point
        self.native_bn = False
        self.new_buffers: Dict[str, torch.Tensor] = {}
        self.acts: Dict[str, torch.Tensor] = {}

    def store(self, t: torch.Tensor, name: Optional[str] = None) -> torch.Tensor:
        """A tensor the CUDA path materialises in HBM in its storage type.
        ``forced_acts`` (name -> tensor read back from ANOTHER evaluation's workspace) replaces the forward VALUE
        by that evaluation's stored one while keeping this graph's derivative: the backward pass then runs, in
        this evaluation's arithmetic, on exactly the forward the other evaluation saw -- a referee for a bf16
        implementation's backward kernels that its forward rounding (ReLU gates, statistics) cannot blur."""
        if self.forced_acts is not None and name is not None and name in self.forced_acts:
            t = t + (self.forced_acts[name].to(t.dtype) - t).detach()
            return _RoundGradBF16.apply(t) if self.emulate_bf16_grads else t
        return _RoundBF16.apply(t) if self.emulate_bf16 else t

    def w(self, name: str) -> torch.Tensor:
        """A conv / transposed-conv weight as the tensor cores see it (bf16 operand in the bf16 mode)."""
        t = self.st[name]
        if name in FP32_OPERANDS:
            return t
        return _RoundBF16Operand.apply(t) if (self.emulate_bf16 or self.forced_acts is not None) else t

    def save(self, name: str, t: torch.Tensor):
        if self.keep:
            self.acts[name] = t

    def relu(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """ReLU (model.py:44,53,75,83,117,184).  With ``relu_masks`` (name -> bool tensor, the sign pattern of
        ANOTHER evaluation's stored activation) the gate is frozen to that pattern: t * mask.  Used to referee
        a bf16 implementation's backward pass at 1e-2: rounding activations to bf16 flips ~1 % of the gates,
        and a flipped gate changes its gradient entry by O(1); with the gates pinned, what is left is the
        arithmetic of the implementation itself."""
        if self.relu_masks is not None and name in self.relu_masks:
            return t * self.relu_masks[name].to(t.dtype)
        return torch.relu(t)


def _batchnorm(ctx: _Ctx, y: torch.Tensor, prefix: str) -> torch.Tensor:
    """nn.BatchNorm2d with default arguments."""
    st = ctx.st
    gamma, beta = st[prefix + ".weight"], st[prefix + ".bias"]
    if ctx.native_bn:
        # the very kernel the reference's nn.BatchNorm2d calls; used for the timed CPU baseline
        return F.batch_norm(y, st[prefix + ".running_mean"], st[prefix + ".running_var"], gamma, beta,
                            ctx.training, BN_MOMENTUM, BN_EPS)
    if ctx.training:
        m = y.shape[0] * y.shape[2] * y.shape[3]
        mean = y.mean(dim=(0, 2, 3))
        var = ((y - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))       # biased
        with torch.no_grad():
            unbiased = var * (m / (m - 1)) if m > 1 else var
            ctx.new_buffers[prefix + ".running_mean"] = (
                (1 - BN_MOMENTUM) * st[prefix + ".running_mean"] + BN_MOMENTUM * mean.detach())
            ctx.new_buffers[prefix + ".running_var"] = (
                (1 - BN_MOMENTUM) * st[prefix + ".running_var"] + BN_MOMENTUM * unbiased.detach())
            ctx.new_buffers[prefix + ".num_batches_tracked"] = st[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = st[prefix + ".running_mean"], st[prefix + ".running_var"]
    xhat = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    return gamma[None, :, None, None] * xhat + beta[None, :, None, None]


def _basic_block(ctx: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    """model.py:39-55 with the 1x1 stride-2 shortcut of model.py:132-138."""
    st = ctx.st
    y1 = ctx.store(F.conv2d(x, ctx.w(p + ".conv1.weight"), stride=2, padding=1), p + ".conv1")
    ctx.save(p + ".conv1", y1)
    a1 = ctx.store(ctx.relu(p + ".relu1", _batchnorm(ctx, y1, p + ".bn1")), p + ".relu1")
    ctx.save(p + ".relu1", a1)
    y2 = ctx.store(F.conv2d(a1, ctx.w(p + ".conv2.weight"), stride=1, padding=1), p + ".conv2")
    ctx.save(p + ".conv2", y2)
    yd = ctx.store(F.conv2d(x, ctx.w(p + ".downsample.0.weight"), stride=2), p + ".downsample.0")
    ctx.save(p + ".downsample.0", yd)
    out = ctx.store(ctx.relu(p, _batchnorm(ctx, y2, p + ".bn2") + _batchnorm(ctx, yd, p + ".downsample.1")), p)
    ctx.save(p, out)
    return out


def _deconv_block(ctx: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    """model.py:70-85 with the upsample branch of model.py:197-204."""
    st = ctx.st
    y1 = ctx.store(F.conv2d(x, ctx.w(p + ".conv1.weight")), p + ".conv1")
    ctx.save(p + ".conv1", y1)
    a1 = ctx.store(ctx.relu(p + ".relu1", _batchnorm(ctx, y1, p + ".bn1")), p + ".relu1")
    ctx.save(p + ".relu1", a1)
    y2 = ctx.store(F.conv_transpose2d(a1, ctx.w(p + ".conv2.weight"), stride=2, padding=1), p + ".conv2")
    ctx.save(p + ".conv2", y2)
    yu = ctx.store(F.conv_transpose2d(x, ctx.w(p + ".upsample.0.weight"), stride=2, padding=1), p + ".upsample.0")
    ctx.save(p + ".upsample.0", yu)
    out = ctx.store(ctx.relu(p, _batchnorm(ctx, y2, p + ".bn2") + _batchnorm(ctx, yu, p + ".upsample.1")), p)
    ctx.save(p, out)
    return out


def encode(ctx: _Ctx, cfg: VAEConfig, x: torch.Tensor):
    st = ctx.st
    y = ctx.store(F.conv2d(x, ctx.w("encoder.conv1.weight"), stride=2, padding=2), "encoder.conv1")    # model.py:115
    ctx.save("encoder.conv1", y)
    a = ctx.store(ctx.relu("encoder.relu", _batchnorm(ctx, y, "encoder.bn1")), "encoder.relu")     # model.py:116-117
    ctx.save("encoder.relu", a)
    for i in range(1, 5):                                                      # model.py:119-122
        a = _basic_block(ctx, a, f"encoder.layer{i}.0")
    pooled = a.mean(dim=(2, 3), keepdim=True)                                  # model.py:123
    mu = F.conv2d(pooled, st["encoder.conv_mu.weight"])                      # model.py:125
    logvar = None
    if cfg.require_rsample:
        logvar = F.conv2d(pooled, st["encoder.conv_logvar.weight"])          # model.py:128
    return mu, logvar


def decode(ctx: _Ctx, cfg: VAEConfig, z: torch.Tensor) -> torch.Tensor:
    st = ctx.st
    zs = ctx.store(z, "decoder.input")
    ctx.save("decoder.input", zs)
    y = ctx.store(F.conv_transpose2d(zs, ctx.w("decoder.conv1.weight")), "decoder.conv1")    # model.py:182
    ctx.save("decoder.conv1", y)
    a = ctx.store(ctx.relu("decoder.relu", _batchnorm(ctx, y, "decoder.bn1")), "decoder.relu")     # model.py:183-184
    ctx.save("decoder.relu", a)
    for i in range(1, len(cfg.dec_planes) + 1):                                # model.py:186-192
        a = _deconv_block(ctx, a, f"decoder.uplayer{i}.0")
    y = ctx.store(F.conv2d(a, ctx.w("decoder.conv2.weight"), st["decoder.conv2.bias"], padding=1), "decoder.conv2")
    ctx.save("decoder.conv2", y)
    out = _batchnorm(ctx, y, "decoder.bn2")                                    # model.py:193
    adj = cfg.adjust
    if adj != 0:                                                               # model.py:328-329
        out = out[:, :, adj:-adj, adj:-adj]
    return out


def forward(st, cfg: VAEConfig, x: torch.Tensor, eps: Optional[torch.Tensor],
            training: bool = True, keep_activations: bool = False, emulate_bf16: bool = False, relu_masks=None,
            forced_acts=None, emulate_bf16_grads: bool = False):
    """VAE.forward (model.py:316-342) for the ``pixelcnn is None`` configuration.
    ``eps`` is the standard-normal draw of ``rsample`` (model.py:149-150),
    shape [N, z, 1, 1].  Returns (mu, logvar, encoding, reconstruction, ctx)."""
    ctx = _Ctx(st, training, keep_activations, emulate_bf16, relu_masks, forced_acts, emulate_bf16_grads)
    mu, logvar = encode(ctx, cfg, x)
    if cfg.require_rsample:
        encoding = mu + eps * torch.exp(0.5 * logvar)
    else:
        encoding = mu
    recon = decode(ctx, cfg, encoding)
    return mu, logvar, encoding, recon, ctx


# ----------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------

def kl_sum(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """model.py:364-365."""
    return -0.5 * torch.sum(logvar - torch.exp(logvar) - mu * mu + 1)


def gaussian_nll_sum(recon: torch.Tensor, target: torch.Tensor, sigma: float) -> torch.Tensor:
    """-Normal(recon, sigma).log_prob(target).sum()  (model.py:403)."""
    return torch.sum((target - recon) ** 2 / (2 * sigma * sigma) + math.log(sigma) + 0.5 * math.log(2 * math.pi))


def weighted_ce_sum(recon: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    """F.cross_entropy(recon, target, reduction='none', weight=w).sum() (model.py:400-401)."""
    logp = recon - torch.logsumexp(recon, dim=1, keepdim=True)
    picked = torch.gather(logp, 1, target[:, None]).squeeze(1)
    if weight is not None:
        picked = picked * weight[target]
    return -picked.sum()


def compute_mmd(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """model.py:367-383: RBF-kernel MMD between ``x`` (true_samples [N, z]) and ``y`` (encoding [N, z]);
    k(a, b) = exp(-mean_d((a_d - b_d)^2) / dim) over all N x N pairs.  VAE.loss returns this / N as its 4th value
    (model.py:394-396,406); with mmd = 0 it is a diagnostic outside the loss."""
    def kern(a, b):
        dim = a.shape[1]
        d2 = ((a[:, None, :] - b[None, :, :]) ** 2).mean(dim=2) / float(dim)
        return torch.exp(-d2)
    return kern(x, x).sum() + kern(y, y).sum() - 2 * kern(x, y).sum()


def loss(cfg: VAEConfig, target, mu, logvar, recon, ce_weight=None, kl_weight: Optional[float] = None):
    """VAE.loss (model.py:385-406) without the MMD diagnostic (coefficient 0 in
    every in-scope model).
    Returns (loss, pxz/N, kl/N) with the last two as 0-d tensors."""
    n = target.shape[0]
    klw = cfg.kl if kl_weight is None else kl_weight
    kl = kl_sum(mu, logvar) if (mu is not None and logvar is not None) else torch.zeros((), dtype=recon.dtype)
    if cfg.categorical:
        pxz = cfg.nll * weighted_ce_sum(recon, target, ce_weight)
    else:
        pxz = cfg.nll * gaussian_nll_sum(recon, target, cfg.sigma_decoder)
    total = (pxz + klw * kl) / n
    return total, pxz.detach() / n, kl.detach() / n


# ----------------------------------------------------------------------------
# one training step (main.py:389-390, main.py:398)
# ----------------------------------------------------------------------------

@dataclass
class StepResult:
    loss: float
    pxz: float
    kl: float
    mu: torch.Tensor
    logvar: Optional[torch.Tensor]
    encoding: torch.Tensor
    recon: torch.Tensor
    grads: Dict[str, torch.Tensor]
    new_buffers: Dict[str, torch.Tensor]
    acts: Dict[str, torch.Tensor] = field(default_factory=dict)


def train_step(st, cfg: VAEConfig, x, target, eps, ce_weight=None, kl_weight=None,
               keep_activations: bool = False, dtype=None, emulate_bf16: bool = False, relu_masks=None,
               forced_acts=None, emulate_bf16_grads: bool = False) -> StepResult:
    """forward -> loss -> backward on the CPU.  ``dtype=torch.float64`` gives a
    higher-precision referee for the 1e-5 fp32 comparison.  ``emulate_bf16`` rounds
    every tensor the CUDA bf16 mode keeps in HBM (conv outputs, activations and their
    gradients) to bf16 at the point it is stored, leaving the arithmetic in ``dtype``:
    the referee for the bf16 mode's *implementation*, as opposed to bf16's own distance
    from fp32."""
    names = [n for n, _ in param_specs(cfg)]
    work = {}
    for k, v in st.items():
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        work[k] = t
    for n in names:
        work[n].requires_grad_(True)
    if dtype is not None:
        x = x.to(dtype)
        eps = eps.to(dtype) if eps is not None else None
        if target.is_floating_point():
            target = target.to(dtype)
        if ce_weight is not None:
            ce_weight = ce_weight.to(dtype)
    mu, logvar, enc, recon, ctx = forward(work, cfg, x, eps, training=True, keep_activations=keep_activations,
                                          emulate_bf16=emulate_bf16, relu_masks=relu_masks, forced_acts=forced_acts,
                                          emulate_bf16_grads=emulate_bf16_grads)
    total, pxz, kl = loss(cfg, target, mu, logvar, recon, ce_weight, kl_weight)
    grads = torch.autograd.grad(total, [work[n] for n in names], allow_unused=True)
    gd = {n: (g if g is not None else torch.zeros_like(work[n])) for n, g in zip(names, grads)}
    return StepResult(float(total.detach()), float(pxz), float(kl), mu.detach(),
                      None if logvar is None else logvar.detach(), enc.detach(), recon.detach(),
                      gd, ctx.new_buffers, {k: v.detach() for k, v in ctx.acts.items()})


def make_timed_step(st, cfg: VAEConfig):
    """CPU baseline for bench.py: the same step as ``train_step`` arranged the way the reference runs it
    (persistent leaf parameters, ``loss.backward()`` into ``.grad``, torch's native batch_norm kernel as
    nn.BatchNorm2d calls it) so that the timing is representative of the reference on these cores.
    Returns step(x, target, eps) -> loss value."""
    names = [n for n, _ in param_specs(cfg)]
    work = {k: v.detach().clone() for k, v in st.items()}
    for n in names:
        work[n].requires_grad_(True)

    def step(x, target, eps):
        ctx = _Ctx(work, True, False)
        ctx.native_bn = True
        mu, logvar = encode(ctx, cfg, x)
        enc = mu + eps * torch.exp(0.5 * logvar) if cfg.require_rsample else mu
        recon = decode(ctx, cfg, enc)
        total, _, _ = loss(cfg, target, mu, logvar, recon)
        for n in names:
            work[n].grad = None
        total.backward()
        return float(total.detach())

    return step


def dp_mean_grads(st, cfg: VAEConfig, shards) -> Dict[str, torch.Tensor]:
    """Data-parallel oracle: the mean over ranks of the
    per-shard gradients (BatchNorm statistics stay per shard).  ``shards`` is a
    list of (x, target, eps)."""
    acc = None
    for (x, t, e) in shards:
        r = train_step(st, cfg, x, t, e)
        if acc is None:
            acc = {k: v.clone() for k, v in r.grads.items()}
        else:
            for k, v in r.grads.items():
                acc[k] += v
    return {k: v / len(shards) for k, v in acc.items()}


# ----------------------------------------------------------------------------
# synthetic Moving-MNIST-like input
# ----------------------------------------------------------------------------

DATA_MEAN = 0.0521   # k=2 label mean  (test-output-models.ipynb:40-43)
DATA_STD = 0.2222


def synthetic_labels(n_frames: int, size: int = 64, seq_len: int = 20, seed: int = 1234) -> torch.Tensor:
    """k=2 k-means label maps of Moving-MNIST-like sequences: two random binary
    blobs of (28/64)*size pixels per sequence bouncing with constant velocity;
    sequences flattened to frames the way movingmnistdataset.py:20-24 does.
    Returns uint8 [n_frames, size, size] with values in {0, 1}."""
    g = torch.Generator().manual_seed(seed)
    d = max(4, (28 * size) // 64)
    n_seq = (n_frames + seq_len - 1) // seq_len
    frames = torch.zeros(n_seq * seq_len, size, size, dtype=torch.uint8)
    for s in range(n_seq):
        for _ in range(2):
            yy, xx = torch.meshgrid(torch.arange(d), torch.arange(d), indexing="ij")
            cy, cx = (d - 1) / 2.0, (d - 1) / 2.0
            r = d * (0.25 + 0.15 * torch.rand((), generator=g).item())
            ring = ((yy - cy) ** 2 + (xx - cx) ** 2).sqrt()
            blob = ((ring < r) & (torch.rand(d, d, generator=g) < 0.45)).to(torch.uint8)
            pos = torch.rand(2, generator=g) * (size - d)
            vel = (torch.rand(2, generator=g) - 0.5) * 8.0
            for t in range(seq_len):
                py, px = int(pos[0].item()), int(pos[1].item())
                f = frames[s * seq_len + t]
                f[py:py + d, px:px + d] |= blob
                pos = pos + vel
                for a in range(2):
                    if pos[a] < 0:
                        pos[a] = -pos[a]; vel[a] = -vel[a]
                    if pos[a] > size - d:
                        pos[a] = 2 * (size - d) - pos[a]; vel[a] = -vel[a]
    return frames[:n_frames]


def normalise(labels: torch.Tensor) -> torch.Tensor:
    """main.py:383-388: (x - data_mean) / data_std on the label map, as [N,1,H,W] fp32."""
    n, h, w = labels.shape
    return (labels.float().view(n, 1, h, w) - DATA_MEAN) / DATA_STD


# ----------------------------------------------------------------------------
# layer-local referee of a backward pass (bf16 mode)
# ----------------------------------------------------------------------------

def local_backward(st, cfg: VAEConfig, x, eps, fwd: Dict[str, torch.Tensor], grd: Dict[str, torch.Tensor],
                   d_recon: torch.Tensor, d_mu: Optional[torch.Tensor], d_logvar: Optional[torch.Tensor],
                   dtype=torch.float64):
    """Every backward kernel of an implementation judged on ITS OWN inputs.

    ``fwd[name]`` / ``grd[name]`` are the forward tensors and gradient tensors another evaluation stored (NCHW; names as
    ``_Ctx.save``: raw conv outputs "<conv>", ReLU outputs "encoder.relu" / "<block>.relu1" / "<block>" /
    "decoder.relu", the latent "decoder.input").  For each layer the adjoint of the reference's forward formula
    (BatchNorm2d in training mode + ReLU gate, model.py:41-55,72-85; conv / transposed conv; pooling + heads + rsample,
    model.py:123-128,148-150) is evaluated in ``dtype`` on the STORED inputs of that layer, so one layer's rounding never
    reaches the next one's expectation.  Returns (param_grads, act_grads): expected parameter gradients by name, and
    expected gradient tensors by activation name with the ReLU gate (or None) under which they are to be compared
    (an implementation may store d(activation) before or after the gate of that activation)."""
    P: Dict[str, torch.Tensor] = {}
    A: Dict[str, Tuple[torch.Tensor, Optional[torch.Tensor]]] = {}
    f = {k: v.to(dtype) for k, v in fwd.items()}
    g = {k: v.to(dtype) for k, v in grd.items()}

    def w_of(name):
        t = st[name].to(dtype)
        return t if name in FP32_OPERANDS else st[name].to(torch.bfloat16).to(dtype)

    def bn_bwd(y_name, prefix, gate_name, dA, out_name=None):
        """dA = gradient wrt the (pre-gate) BatchNorm output sum; returns nothing, fills P and A."""
        y = f[y_name]
        gamma = st[prefix + ".weight"].to(dtype)
        m = y.shape[0] * y.shape[2] * y.shape[3]
        mean = y.mean(dim=(0, 2, 3), keepdim=True)
        var = ((y - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
        rstd = 1.0 / torch.sqrt(var + BN_EPS)
        xhat = (y - mean) * rstd
        gg = dA if gate_name is None else dA * (f[gate_name] > 0).to(dtype)
        dbeta = gg.sum(dim=(0, 2, 3))
        dgamma = (gg * xhat).sum(dim=(0, 2, 3))
        P[prefix + ".bias"] = dbeta
        P[prefix + ".weight"] = dgamma
        dY = gamma[None, :, None, None] * rstd * (gg - dbeta[None, :, None, None] / m - xhat * dgamma[None, :, None, None] / m)
        A[y_name] = (dY, None)

    def conv_bwd(kind, wname, in_name, y_name, stride, padding, bias=None, x_in=None):
        """weight gradient (+ data gradient) of one conv from the stored input and the stored dY."""
        xin = (f[in_name] if x_in is None else x_in.to(dtype)).clone().requires_grad_(in_name is not None)
        w = w_of(wname).clone().requires_grad_(True)
        if kind == "conv":
            y = F.conv2d(xin, w, stride=stride, padding=padding)
        else:
            y = F.conv_transpose2d(xin, w, stride=stride, padding=padding)
        outs = [w] + ([xin] if in_name is not None else [])
        res = torch.autograd.grad(y, outs, g[y_name])
        P[wname] = res[0]
        if bias is not None:
            P[bias] = g[y_name].sum(dim=(0, 2, 3))
        return res[1] if in_name is not None else None

    # ---- decoder tail: d_recon -> BatchNorm (no gate) -> conv 3x3 + bias ----
    ndec = len(cfg.dec_planes)
    last = f"decoder.uplayer{ndec}.0"
    dr = d_recon.to(dtype)
    adj = cfg.adjust
    if adj != 0:
        dr = F.pad(dr, (adj, adj, adj, adj))
    bn_bwd("decoder.conv2", "decoder.bn2", None, dr)
    # decoder.conv2.bias feeds a BatchNorm: true gradient 0 (sum of dY over a normalised channel), SURVEY.md Appendix B.1
    dX = conv_bwd("conv", "decoder.conv2.weight", last, "decoder.conv2", 1, 1)
    A[last] = (dX, f[last] > 0)
    # ---- up-blocks, last to first ----
    for i in range(ndec, 0, -1):
        p = f"decoder.uplayer{i}.0"
        src = f"decoder.uplayer{i - 1}.0" if i > 1 else "decoder.relu"
        dOut = g[p]
        bn_bwd(p + ".conv2", p + ".bn2", p, dOut)
        bn_bwd(p + ".upsample.0", p + ".upsample.1", p, dOut)
        d_a1 = conv_bwd("convT", p + ".conv2.weight", p + ".relu1", p + ".conv2", 2, 1)
        A[p + ".relu1"] = (d_a1, f[p + ".relu1"] > 0)
        bn_bwd(p + ".conv1", p + ".bn1", p + ".relu1", g[p + ".relu1"])
        d_in = conv_bwd("conv", p + ".conv1.weight", src, p + ".conv1", 1, 0)
        d_in = d_in + conv_bwd("convT", p + ".upsample.0.weight", src, p + ".upsample.0", 2, 1)
        A[src] = (d_in, f[src] > 0)
    bn_bwd("decoder.conv1", "decoder.bn1", "decoder.relu", g["decoder.relu"])
    dz = conv_bwd("convT", "decoder.conv1.weight", "decoder.input", "decoder.conv1", 1, 0)
    A["decoder.input"] = (dz, None)
    # ---- rsample + heads + pooling (model.py:123-128,148-150) ----
    feat_name = "encoder.layer4.0"
    feat = f[feat_name].clone().requires_grad_(True)
    wmu = st["encoder.conv_mu.weight"].to(dtype).clone().requires_grad_(True)
    pooled = feat.mean(dim=(2, 3), keepdim=True)
    mu = F.conv2d(pooled, wmu)
    outs, seeds = [], []
    dzs = g["decoder.input"]
    if cfg.require_rsample:
        wlv = st["encoder.conv_logvar.weight"].to(dtype).clone().requires_grad_(True)
        lv = F.conv2d(pooled, wlv)
        z = mu + eps.to(dtype) * torch.exp(0.5 * lv)
        total = (z * dzs).sum()
        if d_mu is not None:
            total = total + (mu * d_mu.to(dtype)).sum() + (lv * d_logvar.to(dtype)).sum()
        r = torch.autograd.grad(total, [feat, wmu, wlv])
        P["encoder.conv_logvar.weight"] = r[2]
    else:
        total = (mu * dzs).sum()
        if d_mu is not None:
            total = total + (mu * d_mu.to(dtype)).sum()
        r = torch.autograd.grad(total, [feat, wmu])
    P["encoder.conv_mu.weight"] = r[1]
    A[feat_name] = (r[0], f[feat_name] > 0)
    # ---- encoder blocks, last to first ----
    for i in range(4, 0, -1):
        p = f"encoder.layer{i}.0"
        src = f"encoder.layer{i - 1}.0" if i > 1 else "encoder.relu"
        dOut = g[p]
        bn_bwd(p + ".conv2", p + ".bn2", p, dOut)
        bn_bwd(p + ".downsample.0", p + ".downsample.1", p, dOut)
        d_a1 = conv_bwd("conv", p + ".conv2.weight", p + ".relu1", p + ".conv2", 1, 1)
        A[p + ".relu1"] = (d_a1, f[p + ".relu1"] > 0)
        bn_bwd(p + ".conv1", p + ".bn1", p + ".relu1", g[p + ".relu1"])
        d_in = conv_bwd("conv", p + ".conv1.weight", src, p + ".conv1", 2, 1)
        d_in = d_in + conv_bwd("conv", p + ".downsample.0.weight", src, p + ".downsample.0", 2, 0)
        A[src] = (d_in, f[src] > 0)
    bn_bwd("encoder.conv1", "encoder.bn1", "encoder.relu", g["encoder.relu"])
    if "encoder.conv1" not in g:
        # an implementation that folds the stem BatchNorm's backward into the stem weight gradient never stores this dY:
        # the weight gradient is then judged on the expected dY, rounded to the storage type the operand has
        g["encoder.conv1"] = A["encoder.conv1"][0].to(torch.bfloat16).to(dtype)
    conv_bwd("conv", "encoder.conv1.weight", None, "encoder.conv1", 2, 2, x_in=x)
    return P, A


# ----------------------------------------------------------------------------
# the step after the hot path: optim.Adam (main.py:468, optimizer.step() main.py:399)
# ----------------------------------------------------------------------------

def adam_update(p, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam defaults restated (amsgrad=False, maximize=False): returns (p', m', v') for the 1-based
    ``step``.  L2 weight decay is added to the gradient; eps sits outside the square root, after the bias correction."""
    if weight_decay != 0.0:
        g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


def train_loop(st, cfg: VAEConfig, x, target, eps_list, lr=1e-3, dtype=None):
    """The loop body main.py:389-399 repeated over ``eps_list`` (one rsample draw per iteration) on a fixed batch:
    forward, loss, backward, Adam.  Returns (losses, final_state) with BatchNorm buffers advanced as training does."""
    cur = {k: (v.detach().clone().to(dtype) if (dtype is not None and v.is_floating_point()) else v.detach().clone())
           for k, v in st.items()}
    names = [n for n, _ in param_specs(cfg)]
    mom = {n: torch.zeros_like(cur[n]) for n in names}
    var = {n: torch.zeros_like(cur[n]) for n in names}
    losses = []
    for it, eps in enumerate(eps_list, start=1):
        r = train_step(cur, cfg, x, target, eps, dtype=dtype)
        losses.append(r.loss)
        for n in names:
            cur[n], mom[n], var[n] = adam_update(cur[n], r.grads[n].to(cur[n].dtype), mom[n], var[n], it, lr=lr)
        for k, b in r.new_buffers.items():
            cur[k] = b
    return losses, cur
