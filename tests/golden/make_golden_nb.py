"""Golden fixtures of the notebook variant from the LIVE reference notebook.

Run in the build container:   python tests/golden/make_golden_nb.py
It executes the class definitions of /root/reference/vae-kl.ipynb code cell 5 (``kl_divergence``,
``VAE_Encoder``, ``VAE_Decoder``) UNMODIFIED, loads seeded weights, runs the loop body of cell 8
(forward, cross-entropy / N + KL / N via torch.distributions, backward) on the CPU in fp32 with an explicit
rsample draw, and stores inputs + outputs under tests/golden/nb_*.npz.  Self-generated fixtures: the
notebook stores none.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import nb_oracle as NB  # noqa: E402
from make_golden import summarize  # noqa: E402

CASES = {
    "nb64_n3": (dict(image_size=64), 3, 1.0),
    "nb128_n2": (dict(image_size=128), 2, 0.35),       # annealed KL weight
}


def notebook_classes():
    nb = json.load(open("/root/reference/vae-kl.ipynb"))
    cells = [c for c in nb["cells"] if c["cell_type"] == "code"]
    src = "".join(cells[5]["source"])
    assert "class VAE_Encoder" in src and "class VAE_Decoder" in src
    ns = {}
    exec("import torch\nfrom torch import nn\nfrom torch.nn import functional as F\n"
         "from torch.distributions import Normal\n" + src, ns)
    return ns


def run_reference(ns, cfg, st, x, y, eps, klw):
    from torch.distributions import Normal
    import torch.nn.functional as F
    enc = ns["VAE_Encoder"](cfg.in_channels, cfg.channels, cfg.z_dimensions)
    dec = ns["VAE_Decoder"](cfg.in_channels, cfg.channels, cfg.z_dimensions)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in st.items() if k.startswith("encoder.")}, strict=True)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in st.items() if k.startswith("decoder.")}, strict=True)
    names = ["encoder." + n for n, _ in enc.named_parameters()] + ["decoder." + n for n, _ in dec.named_parameters()]
    assert names == [n for n, _ in NB.param_specs(cfg)]
    mu, logvar = enc(x)
    torch.manual_seed(777)
    probe = torch.empty(eps.shape).normal_()
    assert torch.equal(probe, eps)
    torch.manual_seed(777)
    encoding = enc.rsample(mu, logvar)
    recon = dec(encoding)
    # loop body, vae-kl.ipynb cell 8
    q_z = Normal(torch.tensor(0.), torch.tensor(1.))
    q_z_x = Normal(mu, (0.5 * logvar).exp())
    pxz = (F.cross_entropy(recon, y, reduction='none') / x.shape[0]).sum()
    kl = (torch.distributions.kl.kl_divergence(q_z_x, q_z) / x.shape[0]).sum()
    loss = pxz + klw * kl
    loss.backward()
    grads = {}
    for n, p in enc.named_parameters():
        grads["encoder." + n] = p.grad.detach().clone()
    for n, p in dec.named_parameters():
        grads["decoder." + n] = p.grad.detach().clone()
    return dict(loss=loss.item(), pxz=pxz.item(), kl=kl.item(), mu=mu.detach(), logvar=logvar.detach(),
                enc=encoding.detach(), recon=recon.detach(), grads=grads)


def main():
    torch.set_num_threads(1)
    ns = notebook_classes()
    for name, (kw, n, klw) in CASES.items():
        cfg = NB.NbConfig(**kw)
        st = NB.init_state(cfg, seed=0)
        x, y = NB.synthetic_batch(cfg, n, seed=1234)
        torch.manual_seed(777)
        eps = torch.empty(n, cfg.z_dimensions, cfg.latent_hw, cfg.latent_hw).normal_()
        ref = run_reference(ns, cfg, st, x, y, eps, klw)
        out = {"cfg/image_size": cfg.image_size, "cfg/channels": cfg.channels, "cfg/z_dimensions": cfg.z_dimensions,
               "cfg/n_classes": cfg.n_classes, "n": n, "kl_weight": np.float64(klw), "seed": 0,
               "y": y.numpy().astype(np.uint8), "eps": eps.numpy(),
               "loss": np.float64(ref["loss"]), "pxz": np.float64(ref["pxz"]), "kl": np.float64(ref["kl"]),
               "mu": ref["mu"].numpy(), "logvar": ref["logvar"].numpy(), "enc": ref["enc"].numpy()}
        summarize("recon", ref["recon"], out)
        for k, v in st.items():
            a = v.double().numpy().ravel()
            out["wfp/" + k] = np.array([a.sum(), np.sqrt((a * a).sum())])
        for k, g in ref["grads"].items():
            summarize("grad/" + k, g, out)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "loss", ref["loss"], "pxz", ref["pxz"], "kl", ref["kl"], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
