"""Run the CUDA path through the public module API and package the result like the oracle's StepResult."""
import types

import torch

import mmvae_b200 as M
from oracle import vae_oracle as O


def build_model(cfg: O.VAEConfig, state, precision, device="cuda"):
    m = M.VAE(in_channels=cfg.in_channels, intermediate_channels=32, decoder_out_channels=cfg.decoder_out_channels,
              pixelcnn_out_channels=0, z_dimension=cfg.z_dimension, pixelcnn=False, only_pixelcnn=False,
              nll=cfg.nll, kl=cfg.kl, mmd=0, require_rsample=cfg.require_rsample, sigma_decoder=cfg.sigma_decoder,
              input_image_size=cfg.input_image_size, precision=precision, width=cfg.width)
    missing = m.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to(device)


def train_step(m, cfg, x, target, eps, ce_weight=None, kl_weight=None):
    """main.py:389-390,398 through the drop-in API, with the rsample draw injected."""
    dev = next(m.parameters()).device
    m.train(True)
    m.zero_grad(set_to_none=True)
    xd, td = x.to(dev), target.to(dev)
    mu, logvar, enc, recon = m(xd, eps=eps.to(dev))
    args = types.SimpleNamespace(data_ratio_of_labels=None if ce_weight is None else ce_weight.to(dev))
    loss, pxz, kl, mmd = m.loss(td, mu, logvar, enc, recon, dev, args, kl_weight=kl_weight)
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters()}
    bufs = {n: b.detach().cpu() for n, b in m.named_buffers()}
    return O.StepResult(float(loss.detach()), pxz, kl, mu.detach().cpu(), logvar.detach().cpu(), enc.detach().cpu(),
                        recon.detach().cpu(), grads, bufs)


def workspace_tensor(m, n, name, training=True):
    """Copy a named NHWC workspace tensor back as an NCHW fp32 CPU tensor."""
    desc, ws, info = m._workspace(n, training)
    off, dims = M._lib.workspace_tensor(desc, name)
    numel = dims[0] * dims[1] * dims[2] * dims[3]
    if m.precision == "fp32":
        t = ws[off:off + 4 * numel].view(torch.float32)
    else:
        t = ws[off:off + 2 * numel].view(torch.bfloat16)
    return t.view(dims).permute(0, 3, 1, 2).float().cpu()


def rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
