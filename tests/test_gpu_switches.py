"""The specialised backward kernels of r02 against the kernels they replaced, on the same inputs: the library's A/B switches
are read once per process, so each setting runs in its own interpreter and the gradients meet here.

  default                         cluster BatchNorm backward, early operand loads, hoisted-coefficient apply, fused stem backward
  MMVAE_NO_BN_CLUSTER ...         the grid-barrier sweep kernel everywhere, operands loaded after the dependency wait, the generic
                                  apply kernel, the stem's dY stored and read back by the plain weight-gradient kernel

Both sides store gradients in bf16 between layers, so they differ by independent roundings (a sum taken in another order moves
a bf16 dY by one ulp here and there), not by more: the flat gradient within 5e-3 (measured 2.3e-3), every tensor within 3e-2 (measured 1.4e-2:
BatchNorm affine gradients are sums with heavy cancellation), the loss bit-identical (same forward)."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from golden_util import Golden
from ours_util import build_model, train_step
g = Golden("base64_n32")
m = build_model(g.cfg, g.state(), "bf16")
res = train_step(m, g.cfg, g.x, g.x, g.eps)
torch.save({"loss": res.loss, "grads": res.grads}, sys.argv[2])
'''

OLD_KERNELS = {"MMVAE_NO_BN_CLUSTER": "1", "MMVAE_BN_LATE_LOADS": "1", "MMVAE_BN_APPLY": "0", "MMVAE_NO_STEM_BWD_FUSE": "1"}


def _run(env_extra, path):
    env = {k: v for k, v in os.environ.items() if not k.startswith("MMVAE_")}
    env.update(env_extra)
    subprocess.run([sys.executable, "-c", CHILD, ROOT, path], check=True, env=env, timeout=300)
    return torch.load(path)


@pytest.mark.gpu
def test_r02_backward_kernels_match_the_ones_they_replaced():
    with tempfile.TemporaryDirectory() as d:
        new = _run({}, os.path.join(d, "new.pt"))
        old = _run(OLD_KERNELS, os.path.join(d, "old.pt"))
    assert new["loss"] == old["loss"]
    names = [k for k in new["grads"] if k != "decoder.conv2.bias"]
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
    per = {k: rel(new["grads"][k], old["grads"][k]) for k in names}
    flat = rel(torch.cat([new["grads"][k].reshape(-1) for k in names]), torch.cat([old["grads"][k].reshape(-1) for k in names]))
    worst = sorted(per.items(), key=lambda kv: -kv[1])[:5]
    print(f"new vs replaced kernels: flat gradient rel-L2 {flat:.2e}, worst tensors {worst}")
    assert flat <= 5e-3, flat          # measured 2.3e-3
    assert worst[0][1] <= 3e-2, worst  # measured 1.4e-2 (encoder.layer1.0.bn1.bias)
