"""Debug: one graph-replayed step vs the same step launched from the host (same weights, frames, noise): relative L2 of every
stored gradient tensor (workspace) and parameter gradient, in backward order."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mmvae_b200 as M
from golden_util import Golden
from ours_util import build_model, workspace_tensor
g = Golden("base64_n32")
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
m = build_model(g.cfg, g.state(), "bf16"); x = g.x.cuda()
step = M.GraphedTrainStep(m, x.shape[0], warmup=1)
step(x); torch.cuda.synchronize()
eps1 = m.last_eps.clone()
m2 = build_model(g.cfg, g.state(), "bf16"); m2.train(True)
mu, lv, enc, rec = m2(x, eps=eps1)
l2, *_ = m2.loss(x, mu, lv, enc, rec, x.device, types.SimpleNamespace(data_ratio_of_labels=None))
l2.backward(); torch.cuda.synchronize()
names = ["decoder.conv2"]
for i in range(5, 0, -1):
    p = f"decoder.uplayer{i}.0"; names += [p, p + ".conv2", p + ".upsample.0", p + ".relu1", p + ".conv1"]
names += ["decoder.relu", "decoder.conv1", "decoder.input"]
for i in range(4, 0, -1):
    p = f"encoder.layer{i}.0"; names += [p, p + ".conv2", p + ".downsample.0", p + ".relu1", p + ".conv1"]
names += ["encoder.relu", "encoder.conv1"]
n = x.shape[0]
for k in names:
    a, b = workspace_tensor(m, n, k + ".grad"), workspace_tensor(m2, n, k + ".grad")
    line = f"{k:36s} grad {rel(a, b):.2e}"
    try:
        a2, b2 = workspace_tensor(m, n, k + ".grad2"), workspace_tensor(m2, n, k + ".grad2")
        line += f"   grad2 {rel(a2, b2):.2e}  sum {rel(a + a2, b + b2):.2e}"
    except M.MMVAEError:
        pass
    f1, f2 = workspace_tensor(m, n, k), workspace_tensor(m2, n, k)
    line += f"   fwd {rel(f1, f2):.1e}"
    print(line)
