"""GPU: the CUDA-graph replay of the training step (mmvae_b200.GraphedTrainStep) against the same step launched
from the host through the public module API, on the same batch, weights and rsample noise."""
import types

import pytest
import torch

import mmvae_b200 as M
from golden_util import Golden
from ours_util import build_model, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_graphed_step_matches_eager(prec):
    g = Golden("base64_n32")
    m = build_model(g.cfg, g.state(), prec)
    x = g.x.cuda()
    step = M.GraphedTrainStep(m, x.shape[0], warmup=1)
    nbt0 = int(m.state_dict()["encoder.bn1.num_batches_tracked"])
    loss, pxz, kl = step(x)
    torch.cuda.synchronize()
    eps1 = m.last_eps.clone()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    loss1 = float(loss)
    assert int(m.state_dict()["encoder.bn1.num_batches_tracked"]) == nbt0 + 1
    # same step from the host with the noise the graph drew
    m2 = build_model(g.cfg, g.state(), prec)
    m2.train(True)
    mu, lv, enc, rec = m2(x, eps=eps1)
    l2, *_ = m2.loss(x, mu, lv, enc, rec, x.device, types.SimpleNamespace(data_ratio_of_labels=None))
    l2.backward()
    torch.cuda.synchronize()
    assert abs(loss1 - float(l2)) <= 1e-5 * abs(float(l2))
    tol = 2e-3 if prec == "bf16" else 1e-4          # fp32 atomics of the weight-gradient reduction reorder sums
    bad = [(n, rel_l2(grads[n], p.grad)) for n, p in m2.named_parameters()
           if n != "decoder.conv2.bias" and rel_l2(grads[n], p.grad) > tol]
    assert not bad, bad[:5]
    # a second replay draws different noise and advances the BatchNorm counters again
    step(x)
    torch.cuda.synchronize()
    assert not torch.equal(m.last_eps, eps1)
    assert int(m.state_dict()["encoder.bn1.num_batches_tracked"]) == nbt0 + 2


def test_graphed_step_from_labels():
    from mmvae_b200 import data as D
    torch.manual_seed(0)
    m = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
              sigma_decoder=0.1, input_image_size=64).cuda()
    labels = D.synthetic_labels(16, 64).pin_memory()
    step = M.GraphedTrainStep(m, 16, warmup=1, from_labels=(D.DATA_MEAN, D.DATA_STD))
    loss, _, _ = step(labels)
    torch.cuda.synchronize()
    want = D.prepare_input(labels.cuda())
    assert torch.equal(step.x, want)
    assert torch.isfinite(loss).item()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


def test_graphed_step_prefetch_pipeline():
    """prefetch(): the next batch is copied host-to-device on a copy stream while a step runs; a step() without
    arguments consumes the staged batch.  Same inputs and inputs order as the direct calls."""
    from mmvae_b200 import data as D
    torch.manual_seed(0)
    m = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
              sigma_decoder=0.1, input_image_size=64, require_rsample=False).cuda()
    batches = [D.synthetic_labels(16, 64, seed=100 + b).pin_memory() for b in range(4)]
    step = M.GraphedTrainStep(m, 16, warmup=1, from_labels=(D.DATA_MEAN, D.DATA_STD))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    direct = []
    for b in batches:
        direct.append(float(step(b)[0]))
    m.load_state_dict(sd)                               # same BatchNorm running buffers; the losses do not depend on them
    piped = []
    step.prefetch(batches[0])
    for i in range(4):
        loss = step()[0]
        if i + 1 < 4:
            step.prefetch(batches[i + 1])
        piped.append(float(loss))
        assert torch.equal(step.labels.cpu(), batches[i])
    assert piped == direct, (piped, direct)


def test_prefetcher_double_buffer():
    """mmvae_b200.data.Prefetcher: staged copies arrive in order and a popped tensor stays valid until the next pop."""
    from mmvae_b200.data import Prefetcher
    host = [torch.full((1 << 20,), b, dtype=torch.uint8).pin_memory() for b in range(6)]
    pre = Prefetcher(torch.device("cuda", 0))
    pre.push(host[0])
    seen = []
    for i in range(6):
        t = pre.pop()
        s = t.to(torch.float32).sum()                   # work on the current stream that reads the popped buffer
        if i + 1 < 6:
            pre.push(host[i + 1])
        seen.append(float(s) / (1 << 20))
    assert seen == [float(b) for b in range(6)], seen
    with pytest.raises(RuntimeError):
        pre.pop()
