"""Golden fixtures for the rows next to the hot path, from the LIVE reference (build container only):

  aux_mmd.npz    the MMD diagnostic VAE.loss returns as its 4th value (model.py:367-383,394-396,406): the reference's
                 `loss()` is run with a seeded global generator, its `torch.randn(N, z)` true_samples draw replayed,
                 and (true_samples, encoding, mmd/N) stored; plus compute_mmd on a second (x, y) pair.
  aux_adam.npz   three iterations of the loop body main.py:389-399 (forward, loss, zero_grad, backward,
                 optim.Adam(lr=1e-3).step()) on the base64_n4 inputs: loss per step, and after the last step a
                 (sum, L2) fingerprint + 257 strided samples of every parameter and the BatchNorm buffers.

Run here, where /root/reference exists:   python tests/golden/make_golden_aux.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from oracle import vae_oracle as O  # noqa: E402
from golden_util import Golden, sample_idx  # noqa: E402

ADAM_STEPS = 3
ADAM_LR = 1e-3


def build_reference(cfg, st):
    import model as ref  # /root/reference/model.py
    m = ref.VAE(in_channels=cfg.in_channels, intermediate_channels=32, decoder_out_channels=cfg.decoder_out_channels,
                pixelcnn_out_channels=0, z_dimension=cfg.z_dimension, pixelcnn=False, only_pixelcnn=False, nll=cfg.nll,
                kl=cfg.kl, mmd=0, require_rsample=True, sigma_decoder=cfg.sigma_decoder,
                input_image_size=cfg.input_image_size)
    m.load_state_dict(st, strict=True)
    return m.train(True)


def main():
    torch.set_num_threads(1)
    g = Golden("base64_n4")
    cfg, st = g.cfg, g.state()
    # ---------------- MMD ----------------
    m = build_reference(cfg, st)
    torch.manual_seed(777)
    mu, logvar, enc, recon = m(g.x)                       # draws eps (= g.eps) from the global generator
    torch.manual_seed(4242)
    probe = torch.randn(g.x.shape[0], cfg.z_dimension)    # what loss() will draw (model.py:395)
    torch.manual_seed(4242)
    loss, pxz, kl, mmd = m.loss(g.x, mu, logvar, enc, recon, torch.device("cpu"), types.SimpleNamespace(data_ratio_of_labels=None))
    gen = torch.Generator().manual_seed(5)
    x2 = torch.randn(33, 48, generator=gen)
    y2 = 0.3 + 1.5 * torch.randn(33, 48, generator=gen)
    out = {"true_samples": probe.numpy(), "encoding": enc.detach().view(-1, cfg.z_dimension).numpy(), "mmd_over_n": np.float64(mmd),
           "x2": x2.numpy(), "y2": y2.numpy(), "mmd2": np.float64(m.compute_mmd(x2, y2).item())}
    np.savez_compressed(os.path.join(HERE, "aux_mmd.npz"), **out)
    print("aux_mmd: mmd/N =", mmd, " compute_mmd(x2,y2) =", out["mmd2"])
    # ---------------- Adam trajectory ----------------
    m = build_reference(cfg, st)
    opt = torch.optim.Adam(list(m.parameters()), lr=ADAM_LR)            # main.py:468
    losses = []
    eps_all = []
    for it in range(ADAM_STEPS):
        torch.manual_seed(1000 + it)
        eps_all.append(torch.empty(g.eps.shape).normal_())
        torch.manual_seed(1000 + it)
        mu, logvar, enc, recon = m(g.x)                                  # main.py:389
        assert torch.equal(enc.detach(), (mu + eps_all[-1] * torch.exp(0.5 * logvar)).detach())
        loss, *_ = m.loss(g.x, mu, logvar, enc, recon, torch.device("cpu"), types.SimpleNamespace(data_ratio_of_labels=None))
        losses.append(loss.item())
        opt.zero_grad()                                                  # main.py:397
        loss.backward()                                                  # main.py:398
        opt.step()                                                       # main.py:399
    out = {"losses": np.array(losses, dtype=np.float64), "eps": torch.stack(eps_all).numpy(), "lr": np.float64(ADAM_LR)}
    for k, v in m.state_dict().items():
        a = v.detach().double().numpy().ravel()
        out["final/" + k + "/fp"] = np.array([a.sum(), np.sqrt((a * a).sum())])
        out["final/" + k + "/samples"] = a[sample_idx(a.size)].astype(np.float64)
    # the optimizer alone: torch.optim.Adam (main.py:468 defaults, and once with weight decay) on a fixed gradient sequence
    gen = torch.Generator().manual_seed(6)
    p0 = torch.randn(1000, generator=gen)
    gs = torch.randn(5, 1000, generator=gen) * torch.logspace(-6, 2, 1000)[None]      # entries from 1e-6 to 1e2
    for tag, wd in (("plain", 0.0), ("wd", 0.01)):
        p = torch.nn.Parameter(p0.clone())
        o = torch.optim.Adam([p], lr=ADAM_LR, weight_decay=wd)
        traj = []
        for it in range(5):
            p.grad = gs[it].clone()
            o.step()
            traj.append(p.detach().clone())
        out[f"opt/{tag}/traj"] = torch.stack(traj).numpy()
        out[f"opt/{tag}/exp_avg"] = o.state[p]["exp_avg"].numpy()
        out[f"opt/{tag}/exp_avg_sq"] = o.state[p]["exp_avg_sq"].numpy()
    out["opt/p0"] = p0.numpy()
    out["opt/grads"] = gs.numpy()
    np.savez_compressed(os.path.join(HERE, "aux_adam.npz"), **out)
    print("aux_adam: losses", losses)


if __name__ == "__main__":
    main()
