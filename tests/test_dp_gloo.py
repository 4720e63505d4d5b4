"""CPU, world_size 2 over gloo: the host side of the data-parallel path (mmvae_b200.parallel) -- frame sharding
and the bucketed gradient averaging over the three backward phases' arena ranges -- against the DP oracle
"mean over ranks of the reference's per-shard gradients" (SURVEY.md 8(e)).  The gradients fed in here come from the
CPU oracle: this test checks the exchange step, the GPU tests check the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vae_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mmvae_b200 as M
        from mmvae_b200 import _lib, parallel as PAR
        cfg = O.VAEConfig(input_image_size=32, z_dimension=16)
        st = O.init_state(cfg, seed=3)
        n_global = 6
        x = O.normalise(O.synthetic_labels(n_global, 32))
        eps = torch.randn(n_global, 16, 1, 1, generator=torch.Generator().manual_seed(11))
        b, e = PAR.shard_bounds(n_global, rank, world)
        res = O.train_step(st, cfg, x[b:e], x[b:e], eps[b:e])                # this rank's shard through the oracle
        desc = _lib.make_desc(e - b, 1, 1, 16, 32)
        table = _lib.param_table(desc)
        flat = torch.zeros(_lib.layout(desc).n_params)
        for name, off, shape in table:
            flat[off:off + res.grads[name].numel()] = res.grads[name].reshape(-1)
        mine = flat.clone()
        sync = PAR.GradSync()
        covered = torch.zeros_like(flat, dtype=torch.bool)
        for ph in (_lib.BWD_DECODER, _lib.BWD_ENC_DEEP, _lib.BWD_ENC_SHALLOW):        # the order backward issues them in
            lo, hi = sync.ranges(desc)[ph]
            sync.reduce_range(flat, lo, hi)
            covered[lo:hi] = True
        sync.finish()
        assert bool(covered.all()), "the three phases must cover the whole gradient arena"
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        want = torch.stack(gathered).mean(0)
        assert torch.allclose(flat, want, rtol=1e-6, atol=1e-6)
        assert sync.bytes_reduced == flat.numel() * 4
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_gradient_averaging_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert all(out.get(r) for r in range(world))


def test_shard_bounds_partition():
    from mmvae_b200 import parallel as PAR
    for n in (1, 7, 256, 2048):
        for w in (1, 2, 3, 8):
            spans = [PAR.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _nb_worker(rank, world, port, out):
    """Notebook variant: NotebookVAE's host-side DP hook (one bucket after the backward call) on CPU-resident
    gradients from the oracle; oracle = mean over ranks of the per-shard gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mmvae_b200 as M
        from mmvae_b200 import parallel as PAR
        from oracle import nb_oracle as NB
        cfg = NB.NbConfig(image_size=64)
        st = NB.init_state(cfg, seed=rank)               # different weights per rank: the broadcast must fix that
        m = M.NotebookVAE(1, 32, 32, image_size=64, precision="fp32")
        m.load_state_dict(st)
        PAR.data_parallel(m)
        st0 = NB.init_state(cfg, seed=0)
        assert all(torch.equal(p.detach(), st0[n]) for n, p in m.named_parameters()), "weights must come from rank 0"
        assert (m._grad_sync is not None) == (world > 1)
        n_global = 4
        x, y = NB.synthetic_batch(cfg, n_global, seed=5)
        eps = torch.randn(n_global, 32, 2, 2, generator=torch.Generator().manual_seed(2))
        b, e = PAR.shard_bounds(n_global, rank, world)
        res = NB.train_step(st0, cfg, x[b:e], y[b:e], eps[b:e])
        flat = torch.zeros(m._n_params)
        for name, off, shape in m._ptable:
            flat[off:off + res.grads[name].numel()] = res.grads[name].reshape(-1)
        mine = flat.clone()
        m._grad_sync.reduce_range(flat, 0, m._n_params)   # what loss_backward() does after the library call
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert torch.allclose(flat, torch.stack(gathered).mean(0), rtol=1e-6, atol=1e-7)
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_notebook_gradient_averaging_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_nb_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert all(out.get(r) for r in range(world))
