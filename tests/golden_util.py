"""Helpers shared by the oracle and GPU parity tests: load a golden fixture
(tests/golden/*.npz, produced from the live reference by make_golden.py) and
compare a training-step result against it."""
import glob
import os

import numpy as np
import torch

from oracle import vae_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if not os.path.basename(p).startswith(("init_", "nb", "aux_")))


def sample_idx(numel, n=257):
    return np.unique(np.linspace(0, numel - 1, n).astype(np.int64))


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        z = self.z
        self.cfg = O.VAEConfig(in_channels=int(z["cfg/in_channels"]),
                               decoder_out_channels=int(z["cfg/decoder_out_channels"]),
                               z_dimension=int(z["cfg/z_dimension"]),
                               input_image_size=int(z["cfg/input_image_size"]),
                               sigma_decoder=float(z["cfg/sigma_decoder"]),
                               nll=float(z["cfg/nll"]), kl=float(z["cfg/kl"]))
        self.labels = torch.from_numpy(z["labels"])
        self.x = O.normalise(self.labels)
        self.eps = torch.from_numpy(z["eps"])
        self.ce_weight = torch.from_numpy(z["ce_weight"]) if "ce_weight" in z.files else None
        self.target = self.labels.long() if self.cfg.categorical else self.x

    def state(self):
        """Regenerate the weights the fixture was made with and check their fingerprint."""
        cfg = self.cfg
        st = O.init_state(cfg, seed=int(self.z["init_seed"]))
        g = torch.Generator().manual_seed(99)
        for k in list(st.keys()):
            if st[k].dim() == 1 and k.endswith(".weight"):
                st[k] = 1.0 + 0.2 * (torch.rand(st[k].shape, generator=g) - 0.5)
            elif st[k].dim() == 1 and k.endswith(".bias") and k != "decoder.conv2.bias":
                st[k] = 0.2 * (torch.rand(st[k].shape, generator=g) - 0.5)
        for k, v in st.items():
            if v.is_floating_point():
                a = v.double().numpy().ravel()
                fp = self.z["wfp/" + k]
                assert abs(a.sum() - fp[0]) <= 1e-9 * max(1.0, abs(fp[0])), f"weight drift in {k}"
                assert abs(np.sqrt((a * a).sum()) - fp[1]) <= 1e-9 * max(1.0, fp[1]), f"weight drift in {k}"
        return st

    # ---- comparisons --------------------------------------------------
    def _check_summary(self, prefix, t, rtol, atol_scale, errs):
        """Compare tensor ``t`` against the stored summary under ``prefix``.
        Tolerance: |diff| <= rtol * (L2 of golden / sqrt(numel)) elementwise-ish,
        expressed as relative L2 on what is stored."""
        z = self.z
        a = t.detach().double().cpu().numpy().ravel()
        l2 = float(z[prefix + "/l2"])
        if prefix + "/full" in z.files:
            ref = z[prefix + "/full"].astype(np.float64).ravel()
            assert ref.size == a.size, f"{prefix}: size {a.size} vs {ref.size}"
            d = np.sqrt(((a - ref) ** 2).sum())
            tol = rtol * l2 + atol_scale
            if not d <= tol:
                errs.append(f"{prefix}: |d|={d:.3e} tol={tol:.3e} (l2={l2:.3e})")
        else:
            idx = sample_idx(a.size)
            ref = z[prefix + "/samples"].astype(np.float64)
            d = np.sqrt(((a[idx] - ref) ** 2).sum())
            rl2 = np.sqrt((ref ** 2).sum())
            tol = rtol * rl2 + atol_scale
            if not d <= tol:
                errs.append(f"{prefix}[samples]: |d|={d:.3e} tol={tol:.3e}")
            n2 = np.sqrt((a * a).sum())
            if not abs(n2 - l2) <= 2 * rtol * l2 + atol_scale:
                errs.append(f"{prefix}[l2]: {n2:.6e} vs {l2:.6e}")

    def check_step(self, res, rtol, zero_grad_atol=None, check_buffers=True, buf_rtol=None):
        """``res``: object with loss, pxz, kl, mu, logvar, encoding, recon, grads (dict), new_buffers (dict).
        ``rtol`` is a relative-L2 tolerance per tensor."""
        z = self.z
        errs = []
        for key in ("loss", "pxz", "kl"):
            ref = float(z[key])
            got = float(getattr(res, key))
            if not abs(got - ref) <= rtol * max(abs(ref), 1e-3):
                errs.append(f"{key}: {got!r} vs {ref!r}")
        for key, t in (("mu", res.mu), ("logvar", res.logvar), ("encoding", res.encoding)):
            ref = torch.from_numpy(z[key]).double()
            d = (t.detach().double().cpu().reshape(ref.shape) - ref).norm().item()
            if not d <= rtol * ref.norm().item() + 1e-7:
                errs.append(f"{key}: |d|={d:.3e} ref={ref.norm().item():.3e}")
        assert tuple(res.recon.shape) == tuple(int(v) for v in z["recon/shape"]), "recon shape"
        self._check_summary("recon", res.recon, rtol, 0.0, errs)
        gmax = max(float(z[k]) for k in z.files if k.startswith("grad/") and k.endswith("/l2"))
        for name, _ in O.param_specs(self.cfg):
            g = res.grads[name]
            if name == "decoder.conv2.bias":
                # feeds a BatchNorm -> true gradient is 0: absolute check
                lim = zero_grad_atol if zero_grad_atol is not None else 1e-6 * gmax
                if not g.abs().max().item() <= max(lim, float(z["grad/" + name + "/l2"]) * 4):
                    errs.append(f"grad/{name}: {g.abs().max().item():.3e} should be ~0 (lim {lim:.3e})")
                continue
            self._check_summary("grad/" + name, g, rtol, 1e-7 * gmax, errs)
        if check_buffers:
            brt = buf_rtol if buf_rtol is not None else rtol
            for prefix, _ in O.bn_names(self.cfg):
                for suffix in (".running_mean", ".running_var"):
                    ref = torch.from_numpy(z["buf/" + prefix + suffix]).double()
                    got = res.new_buffers[prefix + suffix].detach().double().cpu()
                    d = (got - ref).norm().item()
                    if not d <= brt * ref.norm().item() + 1e-6:
                        errs.append(f"buf/{prefix}{suffix}: |d|={d:.3e} ref={ref.norm().item():.3e}")
                nbt = int(res.new_buffers[prefix + ".num_batches_tracked"])
                if nbt != int(z["buf/" + prefix + ".num_batches_tracked"]):
                    errs.append(f"buf/{prefix}.num_batches_tracked: {nbt}")
        assert not errs, f"{self.name}: {len(errs)} mismatches:\n  " + "\n  ".join(errs[:40])
