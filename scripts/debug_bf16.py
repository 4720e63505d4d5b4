"""GPU debug: per-layer forward and per-parameter gradient error of a precision mode against the fp64 oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import Golden
from oracle import vae_oracle as O
from ours_util import build_model, rel_l2, train_step, workspace_tensor

name = sys.argv[1] if len(sys.argv) > 1 else "base64_n4"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
g = Golden(name); st = g.state()
r64 = O.train_step(st, g.cfg, g.x, g.target, g.eps, ce_weight=g.ce_weight, dtype=torch.float64, keep_activations=True)
m = build_model(g.cfg, st, prec)
res = train_step(m, g.cfg, g.x, g.target, g.eps, g.ce_weight)
print("loss", res.loss, r64.loss)
for k, ref in r64.acts.items():
    print(f"act  {k:40s} {rel_l2(workspace_tensor(m, g.x.shape[0], k), ref):.3e}")
print("recon", rel_l2(res.recon, r64.recon), "mu", rel_l2(res.mu, r64.mu))
for n, _ in O.param_specs(g.cfg):
    print(f"grad {n:45s} {rel_l2(res.grads[n], r64.grads[n]):.3e}  |ref|={r64.grads[n].norm().item():.3e}")
