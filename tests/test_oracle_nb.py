"""The notebook-variant oracle (oracle/nb_oracle.py) against fixtures generated from the live reference notebook
(tests/golden/make_golden_nb.py executes vae-kl.ipynb's class definitions unmodified)."""
import pytest
import torch

from oracle import nb_oracle as NB
from nb_util import NB_CASES, NbGolden


def test_fixtures_present():
    assert "nb64_n3" in NB_CASES and "nb128_n2" in NB_CASES


@pytest.mark.parametrize("case", NB_CASES)
def test_oracle_matches_notebook(case):
    g = NbGolden(case)
    st = g.state()
    r = NB.train_step(st, g.cfg, g.x, g.y, g.eps, kl_weight=g.kl_weight, dtype=torch.float32)
    worst = g.check(r.loss, r.pxz, r.kl, r.mu, r.logvar, r.encoding, r.logits, r.grads,
                    tol_loss=1e-6, tol_t=1e-5, tol_g=2e-4)
    assert worst < 2e-4


def test_shapes_and_param_order():
    cfg = NB.NbConfig(image_size=128)
    assert cfg.sizes() == (64, 31, 16, 8, 4) and cfg.out_size == 128
    assert NB.NbConfig(image_size=64).sizes() == (32, 15, 8, 4, 2)
    n = sum(int(torch.tensor(s).prod()) for _, s in NB.param_specs(cfg))
    assert n == 165184          # notebook variant parameter count


def test_fp64_close_to_fp32():
    g = NbGolden("nb64_n3")
    st = g.state()
    a = NB.train_step(st, g.cfg, g.x, g.y, g.eps, dtype=torch.float32)
    b = NB.train_step(st, g.cfg, g.x, g.y, g.eps, dtype=torch.float64)
    assert abs(a.loss - b.loss) <= 1e-5 * abs(b.loss)
    for k in a.grads:
        d = (a.grads[k].double() - b.grads[k]).norm() / b.grads[k].norm()
        assert d < 1e-4, (k, float(d))
