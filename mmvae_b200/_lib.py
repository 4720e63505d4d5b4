"""ctypes binding of libmmvae_b200.so (C ABI declared in include/mmvae.h).

The library is the product: if it is missing the package refuses to import -- there is no
PyTorch / CPU fallback behind these calls.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64,
                    c_void_p, create_string_buffer)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmmvae_b200.so")

PREC_FP32, PREC_BF16 = 0, 1
LOSS_GAUSSIAN, LOSS_CATEGORICAL = 0, 1
BWD_DECODER, BWD_ENC_DEEP, BWD_ENC_SHALLOW, BWD_ALL = 1, 2, 4, 7
BWD_DEFER_JOIN = 8
FLAG_FORCE_SIMT = 1
FLAG_DEFER_LOGITS = 2
ARCH_RESNET, ARCH_NOTEBOOK = 0, 1
ABI_VERSION = 4

EXPORTS = [
    "mmvae_abi_version", "mmvae_last_error", "mmvae_layout", "mmvae_param_entry", "mmvae_bn_entry",
    "mmvae_workspace_tensor", "mmvae_forward", "mmvae_decode", "mmvae_loss_scratch_bytes",
    "mmvae_loss_forward", "mmvae_loss_backward", "mmvae_backward", "mmvae_backward_range",
    "mmvae_philox_normal", "mmvae_adam_step", "mmvae_prepare_input", "mmvae_launch_count",
    "mmvae_conv_entry", "mmvae_selftest_tc", "mmvae_bench_conv", "mmvae_debug_set_trace",
    "mmvae_nb_loss_backward", "mmvae_nb_bench_tail", "mmvae_mmd", "mmvae_mmd_scratch_bytes",
    "mmvae_aux_fence",
]


class Desc(Structure):
    _fields_ = [("struct_size", c_int32), ("batch", c_int32), ("in_channels", c_int32),
                ("out_channels", c_int32), ("z_dim", c_int32), ("image_size", c_int32), ("width", c_int32),
                ("require_rsample", c_int32), ("precision", c_int32), ("training", c_int32),
                ("flags", c_int32), ("arch", c_int32), ("reserved", c_int32 * 4)]


class LayoutInfo(Structure):
    _fields_ = [("n_params", c_int64), ("n_bn_buffers", c_int64), ("n_param_tensors", c_int32),
                ("n_bn", c_int32), ("workspace_bytes", c_int64), ("decoder_size", c_int32), ("crop", c_int32),
                ("train_flops", c_int64)]


class LossArgs(Structure):
    _fields_ = [("struct_size", c_int32), ("kind", c_int32), ("nll", c_float), ("kl", c_float),
                ("sigma", c_float), ("batch", c_int32), ("channels", c_int32), ("height", c_int32),
                ("width", c_int32), ("z_dim", c_int32), ("reserved", c_int32), ("kl_dev", c_void_p)]


class MMVAEError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m mmvae_b200.build` (nvcc, sm_100a). "
            "mmvae_b200 has no PyTorch or CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    lib.mmvae_abi_version.restype = c_int32
    lib.mmvae_last_error.restype = c_char_p
    lib.mmvae_layout.argtypes = [POINTER(Desc), POINTER(LayoutInfo)]
    lib.mmvae_param_entry.argtypes = [POINTER(Desc), c_int32, c_char_p, c_size_t, POINTER(c_int64),
                                      POINTER(c_int32), POINTER(c_int32 * 4)]
    lib.mmvae_bn_entry.argtypes = [POINTER(Desc), c_int32, c_char_p, c_size_t, POINTER(c_int32), POINTER(c_int64)]
    lib.mmvae_workspace_tensor.argtypes = [POINTER(Desc), c_char_p, POINTER(c_int64), POINTER(c_int32 * 4)]
    lib.mmvae_forward.argtypes = [POINTER(Desc), P, P, P, P, P, c_uint64, c_uint64, P, P, P, c_size_t, P, P, P, P, P]
    lib.mmvae_decode.argtypes = [POINTER(Desc), P, P, P, P, P, c_size_t, P, P]
    lib.mmvae_loss_scratch_bytes.restype = c_size_t
    lib.mmvae_loss_forward.argtypes = [POINTER(LossArgs), P, P, P, P, P, P, P, P]
    lib.mmvae_loss_backward.argtypes = [POINTER(LossArgs), P, P, P, P, P, P, P, P, P, P]
    lib.mmvae_backward.argtypes = [POINTER(Desc), P, P, P, c_size_t, P, P, P, P, P, c_int32, P]
    lib.mmvae_nb_loss_backward.argtypes = [POINTER(Desc), P, P, P, P, c_size_t, c_float, P, P, P]
    lib.mmvae_nb_bench_tail.argtypes = [POINTER(Desc), c_int32, P, P, P, c_size_t, P, POINTER(c_int64), POINTER(c_int64), P]
    lib.mmvae_backward_range.argtypes = [POINTER(Desc), c_int32, POINTER(c_int64), POINTER(c_int64)]
    lib.mmvae_philox_normal.argtypes = [c_uint64, c_uint64, P, c_uint64, c_int64, P, P]
    lib.mmvae_adam_step.argtypes = [c_int64, P, P, P, P, c_float, c_float, c_float, c_float, c_float, c_int64,
                                    P, c_float, P]
    lib.mmvae_mmd_scratch_bytes.argtypes = [c_int32]
    lib.mmvae_mmd_scratch_bytes.restype = c_size_t
    lib.mmvae_mmd.argtypes = [P, P, c_int32, c_int32, P, P, P]
    lib.mmvae_aux_fence.argtypes = [P]
    lib.mmvae_prepare_input.argtypes = [P, c_int64, c_float, c_float, P, P, P]
    lib.mmvae_conv_entry.argtypes = [POINTER(Desc), c_int32, c_char_p, c_size_t, POINTER(c_int32 * 8)]
    lib.mmvae_selftest_tc.argtypes = [POINTER(Desc), P, P, c_size_t, P, P, P, c_int32, P]
    lib.mmvae_bench_conv.argtypes = [POINTER(Desc), c_int32, c_int32, P, P, c_size_t, P, POINTER(c_int64), POINTER(c_int64), P]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("mmvae_last_error", "mmvae_loss_scratch_bytes", "mmvae_launch_count", "mmvae_debug_set_trace",
                        "mmvae_mmd_scratch_bytes"):
            fn.restype = c_int32
    lib.mmvae_launch_count.restype = c_int64
    lib.mmvae_debug_set_trace.argtypes = [P]
    lib.mmvae_debug_set_trace.restype = None
    if lib.mmvae_abi_version() != ABI_VERSION:
        raise ImportError(f"libmmvae_b200.so ABI {lib.mmvae_abi_version()} != binding {ABI_VERSION}; rebuild it")
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.mmvae_last_error().decode("utf-8", "replace")
        raise MMVAEError(f"{what} failed ({rc}): {msg}")


def make_desc(batch, in_channels, out_channels, z_dim, image_size, width=1, require_rsample=True,
              precision=PREC_BF16, training=True, flags=0, arch=ARCH_RESNET):
    d = Desc()
    d.struct_size = ctypes.sizeof(Desc)
    d.batch, d.in_channels, d.out_channels, d.z_dim = int(batch), int(in_channels), int(out_channels), int(z_dim)
    d.image_size, d.width = int(image_size), int(width)
    d.require_rsample, d.precision, d.training = int(bool(require_rsample)), int(precision), int(bool(training))
    d.flags = int(flags)
    d.arch = int(arch)
    return d


def layout(desc):
    info = LayoutInfo()
    check(lib.mmvae_layout(byref(desc), byref(info)), "mmvae_layout")
    return info


def param_table(desc):
    """[(name, offset, shape)] in the reference's named_parameters() order."""
    info = layout(desc)
    out = []
    name = create_string_buffer(256)
    off, ndim, shape = c_int64(), c_int32(), (c_int32 * 4)()
    for i in range(info.n_param_tensors):
        check(lib.mmvae_param_entry(byref(desc), i, name, 256, byref(off), byref(ndim), byref(shape)), "mmvae_param_entry")
        out.append((name.value.decode(), off.value, tuple(shape[k] for k in range(ndim.value))))
    return out


def bn_table(desc):
    """[(prefix, channels, buffer_offset)] in state_dict order."""
    info = layout(desc)
    out = []
    name = create_string_buffer(256)
    ch, off = c_int32(), c_int64()
    for i in range(info.n_bn):
        check(lib.mmvae_bn_entry(byref(desc), i, name, 256, byref(ch), byref(off)), "mmvae_bn_entry")
        out.append((name.value.decode(), ch.value, off.value))
    return out


def workspace_tensor(desc, name):
    off, dims = c_int64(), (c_int32 * 4)()
    check(lib.mmvae_workspace_tensor(byref(desc), name.encode(), byref(off), byref(dims)), "mmvae_workspace_tensor")
    return off.value, tuple(dims[k] for k in range(4))


def backward_range(desc, phase):
    b, e = c_int64(), c_int64()
    check(lib.mmvae_backward_range(byref(desc), phase, byref(b), byref(e)), "mmvae_backward_range")
    return b.value, e.value


def conv_table(desc):
    """[(name, kind, k, stride, padding, Ci, Co, H_in, H_out)] in execution order."""
    out = []
    name = create_string_buffer(256)
    shape = (c_int32 * 8)()
    i = 0
    while lib.mmvae_conv_entry(byref(desc), i, name, 256, byref(shape)) == 0:
        out.append((name.value.decode(),) + tuple(shape[k] for k in range(8)))
        i += 1
    return out
