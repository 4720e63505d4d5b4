"""Golden fixtures of the notebook variant (tests/golden/nb_*.npz, made by make_golden_nb.py from the live
reference notebook) and the comparison used by both the oracle test and the GPU parity test."""
import glob
import os

import numpy as np
import torch

from oracle import nb_oracle as NB

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NB_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "nb*.npz")))


def sample_idx(numel, n=257):
    return np.unique(np.linspace(0, numel - 1, n).astype(np.int64))


class NbGolden:
    def __init__(self, name):
        self.name = name
        z = self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.cfg = NB.NbConfig(image_size=int(z["cfg/image_size"]), channels=int(z["cfg/channels"]),
                               z_dimensions=int(z["cfg/z_dimensions"]), n_classes=int(z["cfg/n_classes"]))
        self.n = int(z["n"])
        self.kl_weight = float(z["kl_weight"])
        self.y = torch.from_numpy(z["y"].astype(np.int64))
        self.x = ((self.y.float() / 255.0 - 0.1307) / 0.3081).unsqueeze(1)
        self.eps = torch.from_numpy(z["eps"])

    def state(self):
        st = NB.init_state(self.cfg, seed=int(self.z["seed"]))
        for k, v in st.items():
            a = v.double().numpy().ravel()
            fp = self.z["wfp/" + k]
            assert abs(a.sum() - fp[0]) <= 1e-9 * max(1.0, abs(fp[0])), f"weight drift in {k}"
            assert abs(np.sqrt((a * a).sum()) - fp[1]) <= 1e-9 * max(1.0, fp[1]), f"weight drift in {k}"
        return st

    def rel_err(self, prefix, t):
        """relative L2 distance of tensor t from the stored (full or sampled) golden tensor"""
        z = self.z
        a = t.detach().double().cpu().numpy().ravel()
        if prefix + "/full" in z.files:
            ref = z[prefix + "/full"].astype(np.float64).ravel()
            assert ref.size == a.size, f"{prefix}: size {a.size} vs {ref.size}"
        else:
            ref = z[prefix + "/samples"].astype(np.float64)
            a = a[sample_idx(a.size)]
        return float(np.sqrt(((a - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30))

    def check(self, loss, pxz, kl, mu, logvar, enc, logits, grads, tol_loss, tol_t, tol_g):
        z = self.z
        errs = []
        for nm, v in (("loss", loss), ("pxz", pxz), ("kl", kl)):
            ref = float(z[nm])
            if not abs(v - ref) <= tol_loss * max(abs(ref), 1e-3 if nm == "kl" else 0.0) + (tol_loss if nm == "kl" else 0.0):
                errs.append(f"{nm}: {v} vs {ref}")
        for nm, v in (("mu", mu), ("logvar", logvar), ("enc", enc)):
            ref = torch.from_numpy(z[nm]).double()
            d = float((v.detach().double().cpu() - ref).norm() / ref.norm())
            if not d <= tol_t:
                errs.append(f"{nm}: rel {d:.3e} > {tol_t}")
        if logits is not None:
            d = self.rel_err("recon", logits)
            if not d <= tol_t:
                errs.append(f"recon: rel {d:.3e} > {tol_t}")
        worst = 0.0
        for name, _ in NB.param_specs(self.cfg):
            d = self.rel_err("grad/" + name, grads[name])
            worst = max(worst, d)
            if not d <= tol_g:
                errs.append(f"grad {name}: rel {d:.3e} > {tol_g}")
        assert not errs, f"{self.name}:\n  " + "\n  ".join(errs)
        return worst
