"""CPU: the notebook-variant plan of the C-ABI library (arch = MMVAE_ARCH_NOTEBOOK) agrees with the oracle's parameter
inventory and the host-side module mirrors the notebook's two modules -- no compute."""
import pytest
import torch

from oracle import nb_oracle as NB


@pytest.fixture(scope="module")
def M():
    import mmvae_b200.build as B
    B.build()
    import mmvae_b200
    return mmvae_b200


@pytest.mark.parametrize("size", [64, 128])
def test_layout_matches_oracle_inventory(M, size):
    L = M._lib
    cfg = NB.NbConfig(image_size=size)
    d = L.make_desc(4, 1, cfg.n_classes, cfg.z_dimensions, size, arch=L.ARCH_NOTEBOOK)
    table = L.param_table(d)
    assert [(n, s) for n, _, s in table] == [(n, tuple(s)) for n, s in NB.param_specs(cfg)]
    info = L.layout(d)
    assert info.n_params == 165184 and info.n_bn == 0 and info.decoder_size == size and info.crop == 0
    convs = L.conv_table(d)
    assert [c[0] for c in convs] == ["encoder.conv1", "encoder.conv2", "encoder.conv3", "encoder.conv4", "encoder.conv_mu",
                                     "encoder.conv_logvar", "decoder.conv1", "decoder.conv2", "decoder.conv3", "decoder.conv4"]
    sizes = cfg.sizes()
    assert [c[8] for c in convs[:5]] == list(sizes)                     # H_out of conv1..conv4, conv_mu
    assert [c[7] for c in convs[6:]] == [sizes[-1] * 2, sizes[-1] * 8, sizes[-1] * 16, sizes[-1] * 32]


def test_train_flops_per_frame(M):
    # notebook variant 7714.59 MFLOP/frame at 128x128, 1926.31 at 64x64
    L = M._lib
    for size, mflop in ((128, 7714.59), (64, 1926.31)):
        d = L.make_desc(8, 1, 256, 32, size, arch=L.ARCH_NOTEBOOK)
        assert abs(L.layout(d).train_flops / 8 / 1e6 - mflop) < 0.5, L.layout(d).train_flops / 8 / 1e6


def test_bad_configs_are_refused(M):
    L = M._lib
    for kw in (dict(image_size=96), dict(out_channels=250), dict(z_dim=12)):
        args = dict(batch=2, in_channels=1, out_channels=256, z_dim=32, image_size=64)
        args.update(kw)
        with pytest.raises(M.MMVAEError):
            L.layout(L.make_desc(arch=L.ARCH_NOTEBOOK, **args))


def test_module_mirrors_notebook_modules(M):
    m = M.NotebookVAE(1, 32, 32, image_size=64, precision="fp32")
    names = [n for n, _ in m.named_parameters()]
    assert names == [n for n, _ in NB.param_specs(NB.NbConfig(image_size=64))]
    st = NB.init_state(NB.NbConfig(image_size=64), seed=0)
    m.load_pair({k[8:]: v for k, v in st.items() if k.startswith("encoder.")},
                {k[8:]: v for k, v in st.items() if k.startswith("decoder.")})
    assert torch.equal(m.encoder.conv2.bias.detach(), st["encoder.conv2.bias"])
    # every parameter is a view of the flat arena
    base = m.flat_parameters.data_ptr()
    off = 0
    for p in m.parameters():
        assert p.data_ptr() == base + 4 * off
        off += p.numel()
    # same init calls in the same order as constructing the notebook's modules under the same seed
    torch.manual_seed(5)
    a = M.NotebookVAE(1, 32, 32, image_size=64)
    torch.manual_seed(5)
    enc_w = torch.nn.Conv2d(1, 32, 5, 2, 2)
    assert torch.equal(a.encoder.conv1.weight.detach(), enc_w.weight.detach())
    assert torch.equal(a.encoder.conv1.bias.detach(), enc_w.bias.detach())


def test_no_cpu_fallback(M):
    m = M.NotebookVAE(1, 32, 32, image_size=64)
    with pytest.raises(M.MMVAEError):
        m(torch.zeros(1, 1, 64, 64))
