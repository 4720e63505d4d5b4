"""CPU: the C-ABI library loads, exports every symbol include/mmvae.h declares, its layout agrees with the
oracle's parameter inventory, and the host-side module mirrors the reference interface -- no compute."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def M():
    import mmvae_b200.build as B
    B.build()
    import mmvae_b200
    return mmvae_b200


def test_exports_match_header(M):
    hdr = open(os.path.join(ROOT, "include", "mmvae.h")).read()
    declared = set(re.findall(r"\b(mmvae_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(M._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mmvae.h but not exported"
    assert declared == set(M._lib.EXPORTS)
    assert lib.mmvae_abi_version() == M._lib.ABI_VERSION == 4


def test_struct_sizes(M):
    assert ctypes.sizeof(M._lib.Desc) == 64
    assert ctypes.sizeof(M._lib.LossArgs) == 56


@pytest.mark.parametrize("kw", [dict(), dict(input_image_size=32, z_dimension=32), dict(input_image_size=28, z_dimension=16),
                                dict(decoder_out_channels=2), dict(width=2, z_dimension=256)])
def test_layout_matches_oracle_inventory(M, kw):
    from oracle import vae_oracle as O
    cfg = O.VAEConfig(**kw)
    d = M._lib.make_desc(8, cfg.in_channels, cfg.decoder_out_channels, cfg.z_dimension, cfg.input_image_size, cfg.width)
    table = M._lib.param_table(d)
    specs = O.param_specs(cfg)
    assert [(n, s) for n, _, s in table] == [(n, tuple(s)) for n, s in specs]
    off = 0
    for (n, o, s) in table:
        assert o == off
        off += int(torch.Size(s).numel())
    info = M._lib.layout(d)
    assert info.n_params == off
    assert [(p, c) for p, c, _ in M._lib.bn_table(d)] == O.bn_names(cfg)
    assert info.crop == cfg.adjust and info.decoder_size == (64 if cfg.input_image_size > 32 else 32)


def test_train_flops_per_frame(M):
    # 227.02 MFLOP/frame base, 896.01 widened
    d = M._lib.make_desc(256, 1, 1, 64, 64)
    assert abs(M._lib.layout(d).train_flops / 256 / 1e6 - 227.02) < 0.01
    d = M._lib.make_desc(128, 1, 1, 256, 64, width=2)
    assert abs(M._lib.layout(d).train_flops / 128 / 1e6 - 896.01) < 0.01


def test_bad_desc_is_reported(M):
    d = M._lib.make_desc(8, 1, 1, 64, 128)       # the reference cannot do 128x128 either
    with pytest.raises(M.MMVAEError, match="input_image_size"):
        M._lib.layout(d)
    d = M._lib.make_desc(0, 1, 1, 64, 64)
    with pytest.raises(M.MMVAEError, match="batch"):
        M._lib.layout(d)


def test_backward_ranges_partition_the_arena(M):
    d = M._lib.make_desc(8, 1, 1, 64, 64)
    n = M._lib.layout(d).n_params
    r = [M._lib.backward_range(d, ph) for ph in (M._lib.BWD_ENC_SHALLOW, M._lib.BWD_ENC_DEEP, M._lib.BWD_DECODER)]
    assert r[0][0] == 0 and r[0][1] == r[1][0] and r[1][1] == r[2][0] and r[2][1] == n
    # buckets of 78,240 / 1,181,952 / 852,627 elements
    assert [e - b for b, e in r] == [78240, 1181952, 852627]


def test_module_mirrors_reference_interface(M):
    import inspect
    sig = inspect.signature(M.VAE.__init__)
    names = list(sig.parameters)[1:]
    assert names[:15] == ["in_channels", "intermediate_channels", "decoder_out_channels", "pixelcnn_out_channels",
                          "z_dimension", "pixelcnn", "only_pixelcnn", "pixelcnn_layers", "pixelcnn_activation",
                          "nll", "kl", "mmd", "require_rsample", "sigma_decoder", "input_image_size"]   # model.py:259-262
    assert list(inspect.signature(M.VAE.forward).parameters)[:3] == ["self", "x", "sample"]               # model.py:316
    assert list(inspect.signature(M.VAE.loss).parameters)[:8] == [
        "self", "target", "encoding_mu", "encoding_logvar", "encoding", "reconstruction", "device", "args"]  # model.py:385
    with pytest.raises(NotImplementedError):
        M.VAE(1, 32)                       # defaults are pixelcnn=True, only_pixelcnn=True
    with pytest.raises(NotImplementedError):
        M.VAE(1, 32, pixelcnn=False, only_pixelcnn=False, mmd=1)


def test_state_dict_and_init_match_reference_fingerprint(M):
    """Same seed -> same initial weights as the reference constructor (fingerprint made from the live
    reference by tests/golden/make_init_fingerprint.py)."""
    import numpy as np
    fp = np.load(os.path.join(ROOT, "tests", "golden", "init_fingerprint_seed0.npz"))
    torch.manual_seed(0)
    m = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False,
              only_pixelcnn=False, sigma_decoder=0.1, input_image_size=64)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in fp["keys"]]
    for i, k in enumerate(sd.keys()):
        v = sd[k].double()
        assert tuple(sd[k].shape) == tuple(fp["shapes"][i][: sd[k].dim()])
        assert abs(v.sum().item() - fp["sums"][i]) <= 1e-12 * max(1.0, abs(fp["sums"][i])), k
        assert abs((v * v).sum().item() - fp["sumsq"][i]) <= 1e-12 * max(1.0, fp["sumsq"][i]), k
    # all Parameters are views of one arena
    base = m.flat_parameters.data_ptr()
    for (name, off, shape), p in zip(m._ptable, m.parameters()):
        assert p.data_ptr() == base + 4 * off


def test_no_cpu_fallback(M):
    m = M.VAE(1, 32, pixelcnn=False, only_pixelcnn=False, z_dimension=8, input_image_size=32)
    with pytest.raises(M.MMVAEError, match="no CPU fallback"):
        m(torch.zeros(2, 1, 32, 32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmvae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"
