"""Drop-in `VAE` module: the reference's nn.Module surface over the sm_100a library.

Mirrors the public interface of the reference `VAE` (model.py:258-406) for the configuration the
hot path covers (`pixelcnn=False, only_pixelcnn=False`):

  * constructor keyword arguments of model.py:259-262 (plus keyword-only `precision`, `width`);
  * `forward(x, sample=None) -> (mu, logvar, encoding, reconstruction)`  (model.py:316,342),
    `mu / logvar / encoding` shaped [N, z, 1, 1], reconstruction [N, C, S, S];
  * `loss(target, mu, logvar, encoding, reconstruction, device, args)` -> (loss tensor with a
    grad_fn, pxz/N, KL/N, MMD/N)  (model.py:385-406);
  * `get_z_image`, `get_reconstruction`, `kl_divergence`  (model.py:344-365);
  * parameters / buffers registered under the reference's state_dict names and shapes, fp32, so
    `load_state_dict(reference.state_dict())` works in both directions, and the same
    `torch.manual_seed` gives bit-identical initial weights (same init calls in the same order).

Everything numerical happens in libmmvae_b200.so (include/mmvae.h); this file only owns memory
(one flat fp32 parameter arena that every nn.Parameter is a view of, one flat gradient buffer per
backward, the BatchNorm buffer arena, the activation workspace) and the autograd wiring.
Unsupported constructor combinations raise NotImplementedError -- there is no PyTorch fallback.
"""
import ctypes
import math
from ctypes import byref, c_void_p

import torch
from torch import nn

from . import _lib
from ._lib import lib, check


class _Node(nn.Module):
    """Parameter / buffer holder that reproduces the reference's module tree (names only)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("holder module: the computation runs in libmmvae_b200.so via VAE.forward")


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _stream(dev=None):
    """The current stream of the device the tensors live on (not of whatever device happens to be current)."""
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _VAEForwardFn(torch.autograd.Function):
    """forward -> (mu, logvar, encoding, reconstruction); backward fills every parameter's `.grad`.

    The parameters are NOT autograd inputs of this node: its backward writes the flat gradient arena and binds (or
    accumulates into) each `p.grad` itself, exactly what ~93 AccumulateGrad nodes would do, minus their per-parameter
    host work and minus their stream bookkeeping (an AccumulateGrad node lives on the stream that was current when it was
    created and outlives the step through any tensor of the old graph the caller still holds; replaying it under a CUDA
    graph capture on another stream is an error).  `anchor` is a fresh 0-d leaf per call whose only job is to make the
    outputs require grad.  Consequence: `loss.backward()` behaves as in the reference; `torch.autograd.grad(loss,
    parameters)` and per-parameter hooks do not see these gradients."""

    @staticmethod
    def forward(ctx, module, x, eps, anchor):
        mu, logvar, enc, recon, state = module._run_forward(x, eps, training=True)
        ctx.set_materialize_grads(False)               # unused outputs (encoding) arrive as None, not as a zero-filled tensor
        ctx.module = module
        ctx.state = state
        ctx.save_for_backward(x)
        return mu, logvar, enc, recon

    @staticmethod
    def backward(ctx, d_mu, d_logvar, d_enc, d_recon):
        (x,) = ctx.saved_tensors
        # d_recon None (a loss without a reconstruction term, e.g. kl_divergence(mu, logvar).backward()) goes through
        # as NULL: the C ABI treats it as zero and skips the decoder sweep
        grads = ctx.module._run_backward(ctx.state, x, d_mu, d_logvar, d_enc, d_recon)
        with torch.no_grad():
            for p, g in zip(ctx.module._plist, grads):
                if not p.requires_grad:
                    continue
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad += g                        # accumulate like autograd (main.py zeroes before backward)
        return None, None, None, None


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon, target, mu, logvar, largs, ce_weight, scratch):
        dev = mu.device if recon is None else recon.device
        out = torch.empty(3, dtype=torch.float32, device=dev)          # loss, pxz/N, KL/N
        with torch.cuda.device(dev):
            check(lib.mmvae_loss_forward(byref(largs), _ptr(recon), _ptr(target), _ptr(ce_weight), _ptr(mu),
                                         _ptr(logvar), _ptr(out), _ptr(scratch), _stream(dev)), "mmvae_loss_forward")
        ctx.largs, ctx.ce_weight = largs, ce_weight
        ctx.save_for_backward(recon, target, mu, logvar)
        ctx.mark_non_differentiable(out)
        ctx.set_materialize_grads(False)               # no zero-filled gradient tensor for the diagnostics output
        return out[0], out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        recon, target, mu, logvar = ctx.saved_tensors
        if g_loss is None:                             # only the diagnostics output was used
            return None, None, None, None, None, None, None
        g = g_loss.to(torch.float32).contiguous()
        d_recon = torch.empty_like(recon) if (recon is not None and ctx.needs_input_grad[0]) else None
        need_kl = mu is not None and logvar is not None and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        d_mu = torch.empty_like(mu) if need_kl else None
        d_lv = torch.empty_like(logvar) if need_kl else None
        dev = g.device
        with torch.cuda.device(dev):
            check(lib.mmvae_loss_backward(byref(ctx.largs), _ptr(recon), _ptr(target), _ptr(ctx.ce_weight), _ptr(mu),
                                          _ptr(logvar), _ptr(g), _ptr(d_recon), _ptr(d_mu), _ptr(d_lv), _stream(dev)),
                  "mmvae_loss_backward")
        return d_recon, None, d_mu, d_lv, None, None, None


class VAE(nn.Module):
    def __init__(self, in_channels, intermediate_channels, decoder_out_channels=1, pixelcnn_out_channels=2,
                 z_dimension=32, pixelcnn=True, only_pixelcnn=True, pixelcnn_layers=4, pixelcnn_activation="ReLu",
                 nll=1, kl=1, mmd=0, require_rsample=True, sigma_decoder=0.1, input_image_size=64,
                 *, precision="bf16", width=1):
        super().__init__()
        if only_pixelcnn or pixelcnn:
            raise NotImplementedError(
                "mmvae_b200 implements the plain VAE hot path (pixelcnn=False, only_pixelcnn=False); "
                "PixelCNN / PixelVAE models (model.py:212-255) are out of scope and there is no PyTorch fallback")
        if mmd != 0:
            raise NotImplementedError("MMD-VAE (mmd != 0, model.py:367-383) is out of scope; there is no PyTorch fallback")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        # attributes the reference exposes and main.py reads (main.py:381-382,405,411)
        self.in_channels = in_channels
        self.z_dimensions = z_dimension
        self.decoder_out_channels = decoder_out_channels
        self.pixelcnn_out_channels = pixelcnn_out_channels
        self.num_pixelcnn_layers = pixelcnn_layers
        self.require_rsample = require_rsample
        self.nll, self.kl, self.mmd = nll, kl, mmd
        self.sigma_decoder = sigma_decoder
        self.input_image_size = input_image_size
        self.only_pixelcnn = only_pixelcnn
        self.pixelcnn = None
        self.intermediate_channels = intermediate_channels
        self.precision = precision
        self.width = width
        self._prec = _lib.PREC_BF16 if precision == "bf16" else _lib.PREC_FP32
        self.kernel_flags = 0         # _lib.FLAG_FORCE_SIMT: validation of the tcgen05 kernels against the SIMT ones

        d1 = self._desc(1, True)
        info = _lib.layout(d1)               # validates the configuration (raises MMVAEError otherwise)
        self.adjust = info.crop               # model.py:307-310: (64 - S)//2 or (32 - S)//2; 0 at 64 and 32
        self._decoder_size = info.decoder_size
        self._n_params = info.n_params
        self._ptable = _lib.param_table(d1)
        self._btable = _lib.bn_table(d1)
        self._n_bn_buffers = info.n_bn_buffers

        # ---- memory: flat arenas; Parameters and buffers are views ----
        arena = torch.zeros(self._n_params, dtype=torch.float32)
        bn_arena = torch.zeros(self._n_bn_buffers, dtype=torch.float32)
        counters = torch.zeros(len(self._btable), dtype=torch.int64)
        self._plist = []
        for name, off, shape in self._ptable:
            node, leaf = self._node_for(name)
            p = nn.Parameter(arena[off:off + math.prod(shape)].view(shape))
            node.register_parameter(leaf, p)
            self._plist.append(p)
        for i, (prefix, ch, off) in enumerate(self._btable):
            node, _ = self._node_for(prefix + ".x")
            node.register_buffer("running_mean", bn_arena[off:off + ch])
            node.register_buffer("running_var", bn_arena[off + ch:off + 2 * ch])
            node.register_buffer("num_batches_tracked", counters[i])
        self._arena, self._bn_arena, self._counters = arena, bn_arena, counters
        self._reset_parameters()

        self._ws = {}                 # (N, training) -> (desc, workspace tensor)
        self._gen = 0
        self._philox_seed = None
        self._philox_offset = 0
        self._rng_dev = None          # int64[2] device tensor {seed, offset} while a GraphedTrainStep owns the noise
        self.last_eps = None
        self._loss_scratch = None
        self._mmd_scratch = None
        self._side = None             # torch side stream of the MMD diagnostic
        self._kl_dev = None           # device scalar holding the KL weight while a GraphedTrainStep owns it (annealing)
        self.mmd_diagnostic = True    # compute the 4th return value of loss() like the reference (model.py:394-396)
        self._mmd_join_deferred = False   # set by GraphedTrainStep around its step body
        self._mmd_pending = False
        self.last_true_samples = None
        self._mmd_calls = 0
        self._grad_sync = None        # set by mmvae_b200.parallel.DataParallel
        self.defer_metrics = False    # True: loss() returns 0-d device tensors instead of 3 host floats

    # ------------------------------------------------------------------ structure
    def _desc(self, batch, training):
        return _lib.make_desc(batch, self.in_channels, self.decoder_out_channels, self.z_dimensions,
                              self.input_image_size, self.width, self.require_rsample, self._prec, training,
                              flags=self.kernel_flags)

    def _node_for(self, dotted):
        parts = dotted.split(".")
        node = self
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        return node, parts[-1]

    def _reset_parameters(self):
        """Default init of the reference's layers, drawn in the reference's construction order
        (shortcut conv before the block's own convs: model.py:132-141, model.py:197-208), so that the
        same torch.manual_seed yields the same weights as constructing the reference VAE."""
        by_name = {n: p for (n, _, _), p in zip(self._ptable, self._plist)}

        def conv_init(name, bias=None):
            w = by_name[name]
            t = torch.empty(w.shape, dtype=torch.float32)
            nn.init.kaiming_uniform_(t, a=math.sqrt(5))
            with torch.no_grad():
                w.copy_(t)
            if bias is not None:
                fan_in = w.shape[1] * w.shape[2] * w.shape[3]
                b = torch.empty(by_name[bias].shape, dtype=torch.float32)
                bound = 1 / math.sqrt(fan_in)
                nn.init.uniform_(b, -bound, bound)
                with torch.no_grad():
                    by_name[bias].copy_(b)

        conv_init("encoder.conv1.weight")
        for i in range(1, 5):
            p = f"encoder.layer{i}.0"
            conv_init(p + ".downsample.0.weight")
            conv_init(p + ".conv1.weight")
            conv_init(p + ".conv2.weight")
        conv_init("encoder.conv_mu.weight")
        if self.require_rsample:
            conv_init("encoder.conv_logvar.weight")
        conv_init("decoder.conv1.weight")
        i = 1
        while f"decoder.uplayer{i}.0.conv1.weight" in by_name:
            p = f"decoder.uplayer{i}.0"
            conv_init(p + ".upsample.0.weight")
            conv_init(p + ".conv1.weight")
            conv_init(p + ".conv2.weight")
            i += 1
        conv_init("decoder.conv2.weight", bias="decoder.conv2.bias")
        with torch.no_grad():
            for (name, _, shape), p in zip(self._ptable, self._plist):
                if len(shape) == 1 and name != "decoder.conv2.bias":
                    p.fill_(1.0 if name.endswith(".weight") else 0.0)
            for (prefix, ch, off) in self._btable:
                self._bn_arena[off:off + ch].zero_()
                self._bn_arena[off + ch:off + 2 * ch].fill_(1.0)
            self._counters.zero_()

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._reflatten()
        return self

    def _reflatten(self):
        """Re-establish the flat arenas after .to()/.cuda() replaced the tensors one by one."""
        # nn.Module._apply may have replaced the Parameter objects: re-read them from the tree
        self._plist = []
        for name, _, _ in self._ptable:
            node, leaf = self._node_for(name)
            self._plist.append(node._parameters[leaf])
        dev = self._plist[0].device
        for p in self._plist:
            if p.dtype != torch.float32:
                raise TypeError("mmvae_b200.VAE keeps fp32 master parameters; use precision='bf16' for bf16 compute")
        arena = torch.empty(self._n_params, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for (name, off, shape), p in zip(self._ptable, self._plist):
                n = math.prod(shape)
                arena[off:off + n].copy_(p.data.reshape(-1))
                p.data = arena[off:off + n].view(shape)
            bn_arena = torch.empty(self._n_bn_buffers, dtype=torch.float32, device=dev)
            counters = torch.empty(len(self._btable), dtype=torch.int64, device=dev)
            for i, (prefix, ch, off) in enumerate(self._btable):
                node, _ = self._node_for(prefix + ".x")
                bn_arena[off:off + ch].copy_(node._buffers["running_mean"].to(torch.float32))
                bn_arena[off + ch:off + 2 * ch].copy_(node._buffers["running_var"].to(torch.float32))
                counters[i].copy_(node._buffers["num_batches_tracked"])
                node._buffers["running_mean"] = bn_arena[off:off + ch]
                node._buffers["running_var"] = bn_arena[off + ch:off + 2 * ch]
                node._buffers["num_batches_tracked"] = counters[i]
        self._arena, self._bn_arena, self._counters = arena, bn_arena, counters
        self._ws = {}
        self._loss_scratch = None
        self._mmd_scratch = None
        self._side = None

    def _check_arena(self):
        base = self._arena.data_ptr()
        for (name, off, shape), p in ((self._ptable[0], self._plist[0]), (self._ptable[-1], self._plist[-1])):
            if p.data_ptr() != base + 4 * off:
                self._reflatten()
                return

    @property
    def flat_parameters(self):
        """The fp32 parameter arena every nn.Parameter of this module is a view of."""
        return self._arena

    # ------------------------------------------------------------------ execution
    def _workspace(self, n, training):
        key = (n, bool(training), self.kernel_flags)
        hit = self._ws.get(key)
        if hit is None:
            if len(self._ws) >= 4:
                self._ws.clear()
            desc = self._desc(n, training)
            info = _lib.layout(desc)
            ws = torch.empty(info.workspace_bytes, dtype=torch.uint8, device=self._arena.device)
            hit = (desc, ws, info)
            self._ws[key] = hit
        return hit

    def _require_cuda(self, t):
        if not t.is_cuda:
            raise _lib.MMVAEError(
                "mmvae_b200.VAE runs on a B200 only: move the module and its inputs to CUDA "
                "(there is no CPU fallback)")

    def _run_forward(self, x, eps, training):
        self._require_cuda(x)
        self._require_cuda(self._arena)
        self._check_arena()
        n = x.shape[0]
        z = self.z_dimensions
        desc, ws, info = self._workspace(n, training)
        dev = x.device
        mu = torch.empty(n, z, 1, 1, dtype=torch.float32, device=dev)
        logvar = torch.empty(n, z, 1, 1, dtype=torch.float32, device=dev) if self.require_rsample else None
        enc = torch.empty(n, z, 1, 1, dtype=torch.float32, device=dev)
        d = self._decoder_size
        recon = torch.empty(n, self.decoder_out_channels, d, d, dtype=torch.float32, device=dev)
        eps_out = None
        seed = offset = 0
        rng = None
        if self.require_rsample:
            if eps is None:
                if self._philox_seed is None:
                    self._philox_seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
                if self._rng_dev is not None:
                    # device-resident {seed, offset}: a captured graph draws fresh noise on every replay
                    rng = self._rng_dev
                else:
                    seed, offset = self._philox_seed, self._philox_offset
                    self._philox_offset += (n * z + 3) // 4
                eps_out = torch.empty(n, z, 1, 1, dtype=torch.float32, device=dev)
            else:
                eps = eps.to(device=dev, dtype=torch.float32).contiguous()
                if eps.numel() != n * z:
                    raise ValueError("eps must have N*z elements")
        self._gen += 1
        with torch.cuda.device(dev):
            check(lib.mmvae_forward(byref(desc), _ptr(x), _ptr(self._arena), _ptr(self._bn_arena), _ptr(self._counters),
                                    _ptr(eps), seed, offset, _ptr(eps_out), _ptr(rng), _ptr(ws), ws.numel(), _ptr(mu),
                                    _ptr(logvar), _ptr(enc), _ptr(recon), _stream(dev)), "mmvae_forward")
            # with a device-resident generator state the call advanced rng[1] itself (its heads kernel)
        self.last_eps = eps_out if eps_out is not None else eps
        return mu, logvar, enc, recon, [desc, ws, self._gen, False]

    def _run_backward(self, state, x, d_mu, d_logvar, d_enc, d_recon):
        desc, ws, gen, consumed = state
        if gen != self._gen:
            raise RuntimeError("the activation workspace was overwritten by a later forward(); "
                               "call backward() before the next forward() of the same module")
        if consumed:
            # the BatchNorm-backward accumulators / CTA counters in the workspace are zeroed by the forward only: a second
            # sweep would double-accumulate
            raise RuntimeError("mmvae_b200.VAE: backward() through the same forward() twice (retain_graph) is not "
                               "supported; run forward() again")
        state[3] = True
        if d_recon is not None and tuple(d_recon.shape[-2:]) != (self._decoder_size, self._decoder_size):
            raise RuntimeError("d_recon must be the gradient of the uncropped reconstruction")

        def prep(t):
            return None if t is None else t.to(torch.float32).contiguous()

        d_mu, d_logvar, d_enc, d_recon = prep(d_mu), prep(d_logvar), prep(d_enc), prep(d_recon)
        grads = torch.empty(self._n_params, dtype=torch.float32, device=x.device)
        sync = self._grad_sync
        phases = (_lib.BWD_ALL,) if sync is None else (_lib.BWD_DECODER, _lib.BWD_ENC_DEEP, _lib.BWD_ENC_SHALLOW)
        # a gradient exchange that fences its own stream on the auxiliary streams (GradSync on CUDA) lets the sweep go on
        # without waiting for a phase's weight gradients
        uses_aux = self._prec == _lib.PREC_BF16 and not (self.kernel_flags & _lib.FLAG_FORCE_SIMT)   # who forks auxiliary streams
        defer = _lib.BWD_DEFER_JOIN if (uses_aux and getattr(sync, "fences_aux", False)) else 0
        for ph in phases:
            with torch.cuda.device(x.device):
                flag = defer if ph in (_lib.BWD_DECODER, _lib.BWD_ENC_DEEP) else 0
                check(lib.mmvae_backward(byref(desc), _ptr(x), _ptr(self._arena), _ptr(ws), ws.numel(), _ptr(d_mu),
                                         _ptr(d_logvar), _ptr(d_enc), _ptr(d_recon), _ptr(grads), ph | flag,
                                         _stream(x.device)), "mmvae_backward")
            if sync is not None:
                sync.fence_now = bool(defer) and ph in (_lib.BWD_DECODER, _lib.BWD_ENC_DEEP)
                sync.phase_done(self, desc, grads, ph)
        if sync is not None:
            sync.finish()
        self.last_flat_grad = grads
        return tuple(torch._utils._unflatten_dense_tensors(grads, self._plist))

    def forward(self, x, sample=None, eps=None):
        """model.py:316-342.  `eps` (optional, [N, z, 1, 1]) injects the rsample draw for validation;
        by default it is generated on the device (Philox4x32-10, seeded from torch.initial_seed())
        and kept in `self.last_eps`."""
        self._require_cuda(x)
        x = x.to(torch.float32).contiguous()
        if x.dim() != 4 or x.shape[1] != self.in_channels or x.shape[2] != self.input_image_size:
            raise ValueError(f"expected x of shape [N, {self.in_channels}, {self.input_image_size}, "
                             f"{self.input_image_size}], got {tuple(x.shape)}")
        if self.training and torch.is_grad_enabled():
            anchor = torch.empty((), dtype=torch.float32, device=x.device, requires_grad=True)
            mu, logvar, enc, recon = _VAEForwardFn.apply(self, x, eps, anchor)
        else:
            with torch.no_grad():
                mu, logvar, enc, recon, _ = self._run_forward(x, eps, training=self.training)
        if self.adjust != 0:                                   # model.py:328-329
            a = self.adjust
            recon = recon[:, :, a:-a, a:-a]
        return mu, logvar, enc, recon

    def _decode(self, encoding):
        self._require_cuda(encoding)
        self._check_arena()
        n = encoding.shape[0]
        desc, ws, info = self._workspace(n, self.training)
        e = encoding.detach().to(torch.float32).contiguous()
        d = self._decoder_size
        recon = torch.empty(n, self.decoder_out_channels, d, d, dtype=torch.float32, device=e.device)
        self._gen += 1
        with torch.cuda.device(e.device):
            check(lib.mmvae_decode(byref(desc), _ptr(e), _ptr(self._arena), _ptr(self._bn_arena), _ptr(self._counters),
                                   _ptr(ws), ws.numel(), _ptr(recon), _stream(e.device)), "mmvae_decode")
        if self.adjust != 0:
            a = self.adjust
            recon = recon[:, :, a:-a, a:-a]
        return recon

    def get_z_image(self, encoding):
        """model.py:344-348 (decoder only; no autograd through this path)."""
        return self._decode(encoding)

    def get_reconstruction(self, encoding, sample=None):
        """model.py:353-362 for pixelcnn None."""
        return self._decode(encoding)

    # ------------------------------------------------------------------ loss
    def _loss_args(self, n, c, h, w, nz, kl, nll, kl_dev=None):
        a = _lib.LossArgs()
        a.struct_size = ctypes.sizeof(_lib.LossArgs)
        a.kind = _lib.LOSS_CATEGORICAL if self.decoder_out_channels > self.in_channels else _lib.LOSS_GAUSSIAN
        a.nll, a.kl, a.sigma = float(nll), float(kl), float(self.sigma_decoder)
        a.batch, a.channels, a.height, a.width, a.z_dim = n, c, h, w, nz
        a.kl_dev = kl_dev.data_ptr() if kl_dev is not None else None
        return a

    def _scratch(self, dev):
        if self._loss_scratch is None or self._loss_scratch.device != dev:
            # zero before its first use: it holds the CTA-done counter, which every call leaves at zero again
            self._loss_scratch = torch.zeros(lib.mmvae_loss_scratch_bytes(), dtype=torch.uint8, device=dev)
        return self._loss_scratch

    def compute_mmd(self, true_samples, encoding, out=None):
        """model.py:367-383 divided by N as `loss` returns it (model.py:406): sum k(x,x) + sum k(y,y) - 2 sum k(x,y)
        with k(a,b) = exp(-mean((a-b)^2)/dim), as a 0-d device tensor.  One kernel (mmvae_mmd)."""
        self._require_cuda(encoding)
        dev = encoding.device
        y = encoding.detach().to(torch.float32).reshape(encoding.shape[0], -1).contiguous()
        x = true_samples.to(device=dev, dtype=torch.float32).contiguous()
        if x.shape != y.shape:
            raise ValueError(f"true_samples {tuple(x.shape)} and encoding {tuple(y.shape)} differ in shape")
        n, z = y.shape
        need = lib.mmvae_mmd_scratch_bytes(n)
        if self._mmd_scratch is None or self._mmd_scratch.device != dev or self._mmd_scratch.numel() < need:
            self._mmd_scratch = torch.zeros(need, dtype=torch.uint8, device=dev)
        if out is None:
            out = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mmvae_mmd(_ptr(x), _ptr(y), n, z, _ptr(out), _ptr(self._mmd_scratch), _stream(dev)), "mmvae_mmd")
        return out

    def _mmd_fork(self, encoding, true_samples):
        """Start the MMD diagnostic on a side stream, BESIDE the loss kernel that the caller launches next (it has
        coefficient 0 in every supported model, so nothing but the read of its value depends on it); `_mmd_join` makes the
        current stream wait for it.  true_samples default to a Philox draw (stream 1 of the module's generator; the
        reference draws torch.randn on the host, model.py:395) kept in `last_true_samples`.  Returns the 0-d result."""
        dev = encoding.device
        out = torch.empty((), dtype=torch.float32, device=dev)
        n, z = encoding.shape[0], encoding.numel() // encoding.shape[0]
        cur = torch.cuda.current_stream(dev)
        if true_samples is None:
            ts = torch.empty(n, z, dtype=torch.float32, device=dev)
        else:
            ts = true_samples.to(device=dev, dtype=torch.float32).contiguous()
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.cuda.device(dev):
            if true_samples is None:
                if self._philox_seed is None:
                    self._philox_seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
                # counter range disjoint from call to call; with a device-resident rng state (graph replays) the offset of
                # the rsample stream, which advances every replay, is used instead
                off = (1 << 62) + self._mmd_calls * ((n * z + 3) // 4)
                self._mmd_calls += 1
                check(lib.mmvae_philox_normal(self._philox_seed, off, _ptr(self._rng_dev), 1, n * z, _ptr(ts),
                                              _stream(dev)), "mmvae_philox_normal")
            self.compute_mmd(ts, encoding, out=out)
        self.last_true_samples = ts
        return out

    def _mmd_join(self):
        torch.cuda.current_stream(self._side.device).wait_stream(self._side)

    def kl_divergence(self, encoding_mu, encoding_logvar):
        """model.py:364-365: -0.5 * sum(logvar - exp(logvar) - mu^2 + 1)."""
        self._require_cuda(encoding_mu)
        mu = encoding_mu.to(torch.float32).contiguous()
        lv = encoding_logvar.to(torch.float32).contiguous()
        if lv.numel() != mu.numel():
            raise ValueError("mu and logvar differ in size")
        a = self._loss_args(1, 1, 1, 1, mu.numel(), 1.0, 0.0)
        a.kind = _lib.LOSS_GAUSSIAN
        a.sigma = 1.0
        kl, _ = _LossFn.apply(None, None, mu, lv, a, None, self._scratch(mu.device))
        return kl            # (0 + 1*KL)/1, with a grad_fn

    def loss(self, target, encoding_mu, encoding_logvar, encoding, reconstruction, device, args, kl_weight=None,
             true_samples=None):
        """model.py:385-406.  `kl_weight` (optional) overrides the constructor's KL coefficient for this call
        (annealing).  Returns (loss, pxz/N, KL/N, MMD/N); the last three are host floats like the reference's
        `.item()` calls (one device sync), or 0-d device tensors when `self.defer_metrics`.
        The MMD diagnostic (model.py:394-396; coefficient 0 in every supported model, so it is outside the loss and
        its gradient) is computed from `true_samples` ([N, z]; default: a Philox draw kept in `last_true_samples`)
        when `encoding` is not None and `self.mmd_diagnostic`; otherwise it is 0 as in the reference."""
        self._require_cuda(reconstruction)
        recon = reconstruction.to(torch.float32).contiguous()
        n, c, h, w = recon.shape
        categorical = self.decoder_out_channels > self.in_channels
        if target.shape[0] != n:
            raise ValueError(f"target has {target.shape[0]} frames, reconstruction {n}")
        if categorical:
            tgt = target.to(device=recon.device, dtype=torch.int64).contiguous()
            if tgt.numel() != n * h * w:
                raise ValueError(f"categorical target must have N*H*W = {n * h * w} class indices, got {tuple(target.shape)}")
            cew = getattr(args, "data_ratio_of_labels", None)
            if cew is not None:
                cew = cew.to(device=recon.device, dtype=torch.float32).contiguous()
                if cew.numel() != c:
                    raise ValueError(f"data_ratio_of_labels must have {c} entries")
        else:
            tgt = target.to(device=recon.device, dtype=torch.float32).contiguous()
            if tgt.numel() != recon.numel():
                raise ValueError("target and reconstruction differ in size")
            cew = None
        mu = lv = None
        nz = 0
        if encoding_mu is not None and encoding_logvar is not None:       # model.py:390-391
            mu = encoding_mu.to(torch.float32).contiguous()
            lv = encoding_logvar.to(torch.float32).contiguous()
            nz = mu.numel() // n
        kl_dev = self._kl_dev if kl_weight is None else None
        a = self._loss_args(n, c, h, w, nz, self.kl if kl_weight is None else kl_weight, self.nll, kl_dev)
        mmd = None
        if encoding is not None and self.mmd_diagnostic:                   # model.py:394-396
            mmd = self._mmd_fork(encoding, true_samples)
        loss, out = _LossFn.apply(recon, tgt, mu, lv, a, cew, self._scratch(recon.device))
        if mmd is not None:
            if self._mmd_join_deferred:             # GraphedTrainStep joins the side branch at the END of the step: the
                self._mmd_pending = True            # diagnostic then runs beside the backward instead of in front of it
            else:
                self._mmd_join()
        if self.defer_metrics:
            return loss, out[1], out[2], (mmd if mmd is not None else torch.zeros((), device=recon.device))
        vals = out.tolist()
        return loss, vals[1], vals[2], (float(mmd) if mmd is not None else 0.0)

    def __repr__(self):
        kind = "categorical" if self.decoder_out_channels > self.in_channels else f"normal(sigma={self.sigma_decoder})"
        return (f"mmvae_b200.VAE[{self.precision}] {self.input_image_size}x{self.input_image_size}x{self.in_channels} -> "
                f"z={self.z_dimensions}{' (rsample)' if self.require_rsample else ''} -> "
                f"{self.input_image_size}x{self.input_image_size}x{self.decoder_out_channels}, p(x|z) {kind}, "
                f"{self._n_params} parameters, width x{self.width}")
