"""CPU oracle for the notebook variant of the VAE (BASELINE configs[4]).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mmvae_b200/`` may import this file; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it, as the checker / the thing timed
on the host cores.

What it restates (reference = /root/reference/vae-kl.ipynb, code cell 5 and the loop body of cell 8; the
line numbers are those of the .ipynb JSON file as SURVEY.md cites them):

  * ``VAE_Encoder``  vae-kl.ipynb:122-147  conv(1->C,k5,s2,p2) ReLU, conv(C->C,k5,s2,p1) ReLU,
    conv(C->C,k3,s2,p1) ReLU x2, heads conv_mu / conv_logvar (C->z,k3,s2,p1); every conv has a bias.
  * ``rsample``      vae-kl.ipynb:144-146  mu + eps * exp(0.5 * logvar)
  * ``VAE_Decoder``  vae-kl.ipynb:149-166  nearest upsample x2 -> conv3x3(z->C) ELU, x4 -> conv ELU,
    x2 -> conv ELU, x2 -> conv3x3(C->256) (logits over the 256 grey levels).
  * loop body        vae-kl.ipynb:210-233  loss = sum CE(recon, y)/N + kl_weight * sum KL(q||N(0,1))/N
    (the notebook has kl_weight = 1; "KL-annealed" = a per-step scalar).

The arithmetic is PyTorch's (torch 2.11 CPU kernels here); the oracle is written functionally over a
state dict whose keys are ``encoder.conv1.weight`` ... ``decoder.conv4.bias`` (the notebook keeps two
modules, ``encoder`` and ``decoder``).

Parity pin: the notebook stores no golden vectors; ``tests/golden/make_golden_nb.py`` executes the class
definitions of the notebook cell UNMODIFIED in the build container and stores seeded input/output pairs
under ``tests/golden/nb_*.npz``; ``tests/test_oracle_nb.py`` replays them (self-generated fixtures).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class NbConfig:
    in_channels: int = 1
    channels: int = 32            # intermediate_channels, vae-kl.ipynb cell 6
    z_dimensions: int = 32
    n_classes: int = 256          # literal at vae-kl.ipynb:160
    image_size: int = 128

    def sizes(self) -> Tuple[int, ...]:
        """spatial size after conv1..conv4 and the heads"""
        s = self.image_size
        s1 = (s + 4 - 5) // 2 + 1
        s2 = (s1 + 2 - 5) // 2 + 1
        s3 = (s2 + 2 - 3) // 2 + 1
        s4 = (s3 + 2 - 3) // 2 + 1
        s5 = (s4 + 2 - 3) // 2 + 1
        return s1, s2, s3, s4, s5

    @property
    def latent_hw(self) -> int:
        return self.sizes()[-1]

    @property
    def out_size(self) -> int:
        return self.latent_hw * 32       # x2 x4 x2 x2


def param_specs(cfg: NbConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """``list(encoder.parameters()) + list(decoder.parameters())`` order (vae-kl.ipynb cell 7)."""
    c, z, i = cfg.channels, cfg.z_dimensions, cfg.in_channels
    out = []

    def conv(name, co, ci, k):
        out.append((name + ".weight", (co, ci, k, k)))
        out.append((name + ".bias", (co,)))

    conv("encoder.conv1", c, i, 5)
    conv("encoder.conv2", c, c, 5)
    conv("encoder.conv3", c, c, 3)
    conv("encoder.conv4", c, c, 3)
    conv("encoder.conv_mu", z, c, 3)
    conv("encoder.conv_logvar", z, c, 3)
    conv("decoder.conv1", c, z, 3)
    conv("decoder.conv2", c, c, 3)
    conv("decoder.conv3", c, c, 3)
    conv("decoder.conv4", cfg.n_classes, c, 3)
    return out


def init_state(cfg: NbConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """nn.Conv2d default init (U(+-1/sqrt(fan_in)) for weight and bias) from a seeded CPU generator."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    fan = 1
    for name, shape in param_specs(cfg):
        if name.endswith(".weight"):
            fan = shape[1] * shape[2] * shape[3]
        bound = 1.0 / math.sqrt(fan)
        st[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return st


def _rb(t: torch.Tensor, emulate: bool) -> torch.Tensor:
    """round to bf16 storage (straight-through for autograd) when emulating the bf16 mode"""
    if not emulate:
        return t
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def encode(st, cfg: NbConfig, x: torch.Tensor, emulate_bf16: bool = False):
    e = emulate_bf16
    w = (lambda n: _rb(st[n], e))
    h = _rb(F.relu(F.conv2d(x, w("encoder.conv1.weight"), st["encoder.conv1.bias"], stride=2, padding=2)), e)
    h = _rb(F.relu(F.conv2d(h, w("encoder.conv2.weight"), st["encoder.conv2.bias"], stride=2, padding=1)), e)
    h = _rb(F.relu(F.conv2d(h, w("encoder.conv3.weight"), st["encoder.conv3.bias"], stride=2, padding=1)), e)
    h = _rb(F.relu(F.conv2d(h, w("encoder.conv4.weight"), st["encoder.conv4.bias"], stride=2, padding=1)), e)
    mu = _rb(F.conv2d(h, w("encoder.conv_mu.weight"), st["encoder.conv_mu.bias"], stride=2, padding=1), e)
    logvar = _rb(F.conv2d(h, w("encoder.conv_logvar.weight"), st["encoder.conv_logvar.bias"], stride=2, padding=1), e)
    return mu, logvar


def decode(st, cfg: NbConfig, z: torch.Tensor, emulate_bf16: bool = False) -> torch.Tensor:
    e = emulate_bf16
    w = (lambda n: _rb(st[n], e))
    up = (lambda t, f: F.interpolate(t, scale_factor=f, mode="nearest"))
    h = _rb(F.elu(F.conv2d(up(z, 2), w("decoder.conv1.weight"), st["decoder.conv1.bias"], padding=1)), e)
    h = _rb(F.elu(F.conv2d(up(h, 4), w("decoder.conv2.weight"), st["decoder.conv2.bias"], padding=1)), e)
    h = _rb(F.elu(F.conv2d(up(h, 2), w("decoder.conv3.weight"), st["decoder.conv3.bias"], padding=1)), e)
    return _rb(F.conv2d(up(h, 2), w("decoder.conv4.weight"), st["decoder.conv4.bias"], padding=1), e)


def kl_sum(mu, logvar):
    # KL(N(mu, exp(logvar/2)) || N(0,1)) summed, vae-kl.ipynb:225-227 == vae-kl.ipynb:119-120
    return -0.5 * torch.sum(logvar - logvar.exp() - mu * mu + 1)


@dataclass
class NbStep:
    loss: float
    pxz: float
    kl: float
    mu: torch.Tensor
    logvar: torch.Tensor
    encoding: torch.Tensor
    logits: torch.Tensor
    grads: Dict[str, torch.Tensor]


def train_step(st, cfg: NbConfig, x, y, eps, kl_weight: float = 1.0, dtype=torch.float32,
               emulate_bf16: bool = False, keep_logits: bool = True) -> NbStep:
    """One loop body of vae-kl.ipynb:210-233 without the optimizer: x [N,1,S,S] float, y [N,S',S'] int64
    grey levels (S' = decoder output size), eps [N,z,h,w] the rsample draw."""
    p = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in st.items()}
    x = x.to(dtype)
    mu, logvar = encode(p, cfg, x, emulate_bf16)
    enc = mu + eps.to(dtype) * torch.exp(0.5 * logvar)
    logits = decode(p, cfg, _rb(enc, emulate_bf16), emulate_bf16)
    n = x.shape[0]
    pxz = (F.cross_entropy(logits, y, reduction="none") / n).sum()
    kl = kl_sum(mu, logvar) / n
    loss = pxz + kl_weight * kl
    loss.backward()
    return NbStep(loss=float(loss.detach()), pxz=float(pxz.detach()), kl=float(kl.detach()), mu=mu.detach(), logvar=logvar.detach(),
                  encoding=enc.detach(), logits=logits.detach() if keep_logits else None,
                  grads={k: v.grad.detach() for k, v in p.items()})


def make_timed_step(st, cfg: NbConfig):
    """fp32 CPU loop body for bench.py's CPU baseline (forward + loss + backward, fresh leaves each call)."""
    p = {k: v.detach().float().clone().requires_grad_(True) for k, v in st.items()}

    def step(x, y, eps):
        for v in p.values():
            v.grad = None
        mu, logvar = encode(p, cfg, x)
        enc = mu + eps * torch.exp(0.5 * logvar)
        logits = decode(p, cfg, enc)
        n = x.shape[0]
        loss = (F.cross_entropy(logits, y, reduction="none") / n).sum() + kl_sum(mu, logvar) / n
        loss.backward()
        return float(loss.detach())

    return step


def synthetic_batch(cfg: NbConfig, n: int, seed: int = 1234):
    """Moving-MNIST-like frames as the notebook feeds them: grey levels 0..255 (y, int64) and the
    normalised float input x = (y/255 - 0.1307)/0.3081 (vae-kl.ipynb cell 2, cell 8 ``y*255``)."""
    g = torch.Generator().manual_seed(seed)
    s = cfg.image_size
    y = torch.zeros(n, s, s, dtype=torch.int64)
    d = max(4, (28 * s) // 64)
    for i in range(n):
        for _ in range(2):
            oy = int(torch.randint(0, s - d + 1, (1,), generator=g))
            ox = int(torch.randint(0, s - d + 1, (1,), generator=g))
            blob = (torch.rand(d, d, generator=g) * 255).long()
            yy, xx = torch.meshgrid(torch.arange(d), torch.arange(d), indexing="ij")
            mask = ((yy - d / 2) ** 2 + (xx - d / 2) ** 2) < (d / 2.2) ** 2
            y[i, oy:oy + d, ox:ox + d] = torch.where(mask, blob, y[i, oy:oy + d, ox:ox + d])
    x = ((y.float() / 255.0 - 0.1307) / 0.3081).unsqueeze(1)
    return x, y
