"""CPU: pin the oracle's restatements of the rows NEXT to the hot path against fixtures made from the live reference
(tests/golden/make_golden_aux.py): the MMD diagnostic (model.py:367-383,394-396) and the Adam loop body
(main.py:389-399, optim.Adam main.py:468)."""
import os

import numpy as np
import torch

from golden_util import GOLDEN_DIR, Golden, sample_idx
from oracle import vae_oracle as O


def test_oracle_mmd_matches_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "aux_mmd.npz"))
    ts, enc = torch.from_numpy(z["true_samples"]), torch.from_numpy(z["encoding"])
    got = O.compute_mmd(ts, enc).item() / ts.shape[0]
    assert abs(got - float(z["mmd_over_n"])) <= 1e-5 * abs(float(z["mmd_over_n"]))
    got2 = O.compute_mmd(torch.from_numpy(z["x2"]), torch.from_numpy(z["y2"])).item()
    assert abs(got2 - float(z["mmd2"])) <= 1e-5 * abs(float(z["mmd2"]))
    # fp64 evaluation agrees too (the GPU kernel accumulates in fp64)
    got3 = O.compute_mmd(torch.from_numpy(z["x2"]).double(), torch.from_numpy(z["y2"]).double()).item()
    assert abs(got3 - float(z["mmd2"])) <= 1e-5 * abs(float(z["mmd2"]))


def test_oracle_adam_loop_matches_reference():
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLDEN_DIR, "aux_adam.npz"))
    g = Golden("base64_n4")
    eps_list = [torch.from_numpy(e) for e in z["eps"]]
    losses, final = O.train_loop(g.state(), g.cfg, g.x, g.x, eps_list, lr=float(z["lr"]))
    ref = z["losses"]
    # step 1 is one forward; later losses went through Adam's g/sqrt(v) normalisation, which turns fp32 round-off in
    # tiny gradient entries into O(lr) parameter differences
    assert abs(losses[0] - ref[0]) <= 1e-5 * abs(ref[0])
    for a, b in zip(losses[1:], ref[1:]):
        assert abs(a - b) <= 2e-3 * abs(b), (losses, ref)
    check_final_state(final, g.state(), z, float(z["lr"]), len(eps_list))


def test_oracle_adam_update_matches_torch_optim():
    """The optimizer alone on a fixed gradient sequence (entries spanning 1e-6 .. 1e2): five steps of torch.optim.Adam
    as recorded from the reference's call (main.py:468), with and without weight decay."""
    z = np.load(os.path.join(GOLDEN_DIR, "aux_adam.npz"))
    for tag, wd in (("plain", 0.0), ("wd", 0.01)):
        p = torch.from_numpy(z["opt/p0"]).clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for it in range(5):
            p, m, v = O.adam_update(p, torch.from_numpy(z["opt/grads"][it]), m, v, it + 1, lr=float(z["lr"]), weight_decay=wd)
            want = torch.from_numpy(z[f"opt/{tag}/traj"][it])
            assert (p - want).abs().max().item() <= 2e-6, (tag, it, (p - want).abs().max().item())
        assert torch.allclose(m, torch.from_numpy(z[f"opt/{tag}/exp_avg"]), rtol=2e-4, atol=1e-9)
        assert torch.allclose(v, torch.from_numpy(z[f"opt/{tag}/exp_avg_sq"]), rtol=2e-4, atol=1e-9)


def check_final_state(final, initial, z, lr, steps, upd_tol=None):
    """Parameters after `steps` Adam steps against the fixture's strided samples.  Adam moves every entry by about lr
    per step whatever the gradient's size, so a near-zero gradient entry whose sign differs by round-off lands up to
    2*lr*steps away (27 % of a deep layer's whole update at N=4 between two fp32 evaluations of the same formulas):
    bound the worst entry by that; `upd_tol` additionally bounds the UPDATE (final - initial) by relative L2."""
    for k, v in final.items():
        a = v.detach().double().cpu().numpy().ravel()
        idx = sample_idx(a.size)
        want, got = z["final/" + k + "/samples"], a[idx]
        if k.endswith("num_batches_tracked"):
            assert np.array_equal(got, want), k
            continue
        assert np.abs(got - want).max() <= 2 * lr * steps + 1e-6 + 1e-3 * np.abs(want).max(), (k, np.abs(got - want).max())
        init = initial[k].detach().double().cpu().numpy().ravel()[idx]
        upd = np.linalg.norm(want - init)
        if upd_tol is not None and upd > 0 and k != "decoder.conv2.bias":
            assert np.linalg.norm(got - want) <= upd_tol * upd, (k, np.linalg.norm(got - want) / upd)
