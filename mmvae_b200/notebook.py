"""The notebook variant of the VAE (reference vae-kl.ipynb, code cells 5-8) over the sm_100a library.

The notebook keeps two modules, ``encoder = VAE_Encoder(in_channels, intermediate_channels, z_dimensions)`` and
``decoder = VAE_Decoder(...)`` (vae-kl.ipynb:122-166), and a hand-written loop body (vae-kl.ipynb:210-233):

    mu, logvar = encoder(x); encoding = encoder.rsample(mu, logvar); reconstruction = decoder(encoding)
    px_given_z = (F.cross_entropy(reconstruction, y, reduction='none') / N).sum()
    kl = (kl_divergence(Normal(mu, exp(logvar/2)), Normal(0, 1)) / N).sum()
    loss = px_given_z + kl;  optimizer.zero_grad();  loss.backward()

``NotebookVAE`` holds both parameter sets under the same names (``encoder.conv1.weight`` ...
``decoder.conv4.bias``; ``encoder.state_dict()`` / ``decoder.state_dict()`` interchange with the notebook's
modules) and exposes that loop body as two calls:

    mu, logvar, encoding, reconstruction = model(x)            # forward; reconstruction = logits [N,256,S,S]
    loss, pxz, kl = model.loss_backward(y, kl_weight=1.0)      # CE + KL and every parameter's .grad

At training batch sizes the 256-channel logits are 8 MB per 128x128 frame in bf16: pass ``materialize=False`` to
``forward`` and they stay in the library's workspace (the loss / backward call reads them there).  Everything
numerical runs in libmmvae_b200.so (``arch = MMVAE_ARCH_NOTEBOOK``, include/mmvae.h); there is no PyTorch fallback.
"""
import math
from ctypes import byref, c_void_p

import torch
from torch import nn

from . import _lib
from ._lib import lib, check
from .model import _Node, _ptr, _stream


class NotebookVAE(nn.Module):
    def __init__(self, in_channels=1, intermediate_channels=32, z_dimensions=32, *, image_size=128, n_classes=256,
                 precision="bf16"):
        super().__init__()
        if in_channels != 1:
            raise NotImplementedError("the notebook variant is built for 1-channel frames (vae-kl.ipynb cell 6)")
        if intermediate_channels % 32 != 0:
            raise NotImplementedError("intermediate_channels must be a multiple of 32")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.in_channels, self.intermediate_channels, self.z_dimensions = in_channels, intermediate_channels, z_dimensions
        self.image_size, self.n_classes, self.precision = image_size, n_classes, precision
        self._prec = _lib.PREC_BF16 if precision == "bf16" else _lib.PREC_FP32
        self.kernel_flags = 0
        d1 = self._desc(1)
        info = _lib.layout(d1)
        self._n_params = info.n_params
        self._ptable = _lib.param_table(d1)
        self.latent_hw = image_size // 32
        arena = torch.zeros(self._n_params, dtype=torch.float32)
        self.encoder, self.decoder = _Node(), _Node()
        self._plist = []
        for name, off, shape in self._ptable:
            node, leaf = self._node_for(name)
            p = nn.Parameter(arena[off:off + math.prod(shape)].view(shape))
            node.register_parameter(leaf, p)
            self._plist.append(p)
        self._arena = arena
        self._grads = None
        self._ws = {}
        self._state = None
        self._philox_seed = None
        self._philox_offset = 0
        self.last_eps = None
        self._grad_sync = None        # set by mmvae_b200.parallel.data_parallel
        self._reset_parameters()

    def _desc(self, batch, defer_logits=False):
        flags = self.kernel_flags | (_lib.FLAG_DEFER_LOGITS if defer_logits else 0)
        return _lib.make_desc(batch, 1, self.n_classes, self.z_dimensions, self.image_size,
                              self.intermediate_channels // 32, True, self._prec, True, flags=flags,
                              arch=_lib.ARCH_NOTEBOOK)

    def _node_for(self, dotted):
        parts = dotted.split(".")
        node = self
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        return node, parts[-1]

    def _reset_parameters(self):
        """nn.Conv2d default init in the notebook's construction order (encoder then decoder, weight then bias)."""
        by_name = {n: p for (n, _, _), p in zip(self._ptable, self._plist)}
        for name, _, shape in self._ptable:
            if not name.endswith(".weight"):
                continue
            w = torch.empty(shape, dtype=torch.float32)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            bound = 1 / math.sqrt(shape[1] * shape[2] * shape[3])
            b = torch.empty(shape[0], dtype=torch.float32)
            nn.init.uniform_(b, -bound, bound)
            with torch.no_grad():
                by_name[name].copy_(w)
                by_name[name[:-len("weight")] + "bias"].copy_(b)

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._reflatten()
        return self

    def _reflatten(self):
        self._plist = []
        for name, _, _ in self._ptable:
            node, leaf = self._node_for(name)
            self._plist.append(node._parameters[leaf])
        dev = self._plist[0].device
        arena = torch.empty(self._n_params, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for (name, off, shape), p in zip(self._ptable, self._plist):
                if p.dtype != torch.float32:
                    raise TypeError("NotebookVAE keeps fp32 master parameters; use precision='bf16' for bf16 compute")
                n = math.prod(shape)
                arena[off:off + n].copy_(p.data.reshape(-1))
                p.data = arena[off:off + n].view(shape)
        self._arena = arena
        self._grads = None
        self._ws = {}
        self._state = None

    def _check_arena(self):
        base = self._arena.data_ptr()
        for (name, off, shape), p in ((self._ptable[0], self._plist[0]), (self._ptable[-1], self._plist[-1])):
            if p.data_ptr() != base + 4 * off:
                self._reflatten()
                return

    @property
    def flat_parameters(self):
        return self._arena

    @property
    def flat_grads(self):
        return self._grads

    last_flat_grad = flat_grads                          # the name mmvae_b200.FusedAdam looks for

    def load_pair(self, encoder_state, decoder_state):
        """load the notebook's two state dicts (``encoder.state_dict()``, ``decoder.state_dict()``)"""
        st = {"encoder." + k: v for k, v in encoder_state.items()}
        st.update({"decoder." + k: v for k, v in decoder_state.items()})
        return self.load_state_dict(st)

    def _workspace(self, n, defer_logits=False):
        key = (n, self.kernel_flags, defer_logits)
        hit = self._ws.get(key)
        if hit is None:
            self._ws.clear()
            desc = self._desc(n, defer_logits)
            info = _lib.layout(desc)
            ws = torch.empty(info.workspace_bytes, dtype=torch.uint8, device=self._arena.device)
            hit = (desc, ws, info)
            self._ws[key] = hit
        return hit

    def _require_cuda(self, t):
        if not t.is_cuda:
            raise _lib.MMVAEError("mmvae_b200.NotebookVAE runs on a B200 only: move the module and its inputs to CUDA "
                                  "(there is no CPU fallback)")

    # ------------------------------------------------------------------ the loop body
    @torch.no_grad()
    def forward(self, x, eps=None, materialize=True):
        """encoder -> rsample -> decoder (vae-kl.ipynb:213-215).  x [N,1,S,S] fp32; eps: optional rsample draw
        [N,z,h,h] (default: device Philox stream).  Returns (mu, logvar, encoding, reconstruction); reconstruction
        is None when ``materialize`` is False."""
        self._require_cuda(x)
        self._require_cuda(self._arena)
        self._check_arena()
        x = x.to(torch.float32).contiguous()
        n, z, h = x.shape[0], self.z_dimensions, self.latent_hw
        if tuple(x.shape[1:]) != (1, self.image_size, self.image_size):
            raise ValueError(f"x must be [N,1,{self.image_size},{self.image_size}]")
        # without materialised logits the last conv runs fused with the cross-entropy inside loss_backward()
        desc, ws, info = self._workspace(n, defer_logits=not materialize and self.precision == "bf16")
        dev = x.device
        mu = torch.empty(n, z, h, h, dtype=torch.float32, device=dev)
        logvar = torch.empty_like(mu)
        enc = torch.empty_like(mu)
        recon = torch.empty(n, self.n_classes, self.image_size, self.image_size, dtype=torch.float32, device=dev) if materialize else None
        eps_out, seed, offset = None, 0, 0
        if eps is None:
            if self._philox_seed is None:
                self._philox_seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
            seed, offset = self._philox_seed, self._philox_offset
            self._philox_offset += (mu.numel() + 3) // 4
            eps_out = torch.empty_like(mu)
        else:
            eps = eps.to(device=dev, dtype=torch.float32).contiguous()
            if eps.numel() != mu.numel():
                raise ValueError("eps must have N*z*h*h elements")
        check(lib.mmvae_forward(byref(desc), _ptr(x), _ptr(self._arena), c_void_p(0), c_void_p(0), _ptr(eps), seed, offset,
                                _ptr(eps_out), c_void_p(0), _ptr(ws), ws.numel(), _ptr(mu), _ptr(logvar), _ptr(enc),
                                _ptr(recon), _stream()), "mmvae_forward")
        self.last_eps = eps_out if eps_out is not None else eps
        self._state = (desc, ws, x)
        return mu, logvar, enc, recon

    @torch.no_grad()
    def loss_backward(self, y, kl_weight=1.0):
        """Rest of the loop body (vae-kl.ipynb:225-231) for the batch of the last forward(): returns a device tensor
        [loss, px_given_z, kl] and leaves d loss / d parameter in every parameter's ``.grad`` (overwritten, like
        ``optimizer.zero_grad(); loss.backward()``)."""
        if self._state is None:
            raise RuntimeError("loss_backward() needs a forward() first")
        desc, ws, x = self._state
        y = y.to(device=x.device, dtype=torch.int64).contiguous()
        if tuple(y.shape) != (x.shape[0], self.image_size, self.image_size):
            raise ValueError(f"y must be [N,{self.image_size},{self.image_size}] class indices")
        if self._grads is None or self._grads.device != x.device:
            self._grads = torch.empty(self._n_params, dtype=torch.float32, device=x.device)
            for (name, off, shape), p in zip(self._ptable, self._plist):
                p.grad = self._grads[off:off + math.prod(shape)].view(shape)
        out = torch.empty(3, dtype=torch.float32, device=x.device)
        check(lib.mmvae_nb_loss_backward(byref(desc), _ptr(x), _ptr(y), _ptr(self._arena), _ptr(ws), ws.numel(),
                                         float(kl_weight), _ptr(out), _ptr(self._grads), _stream()),
              "mmvae_nb_loss_backward")
        if self._grad_sync is not None:                 # data parallel: mean over ranks of the per-shard gradients
            self._grad_sync.reduce_range(self._grads, 0, self._n_params)
        return out

    DATA_MEAN, DATA_STD = 0.1307, 0.3081               # transforms.Normalize of vae-kl.ipynb cell 2

    @torch.no_grad()
    def prepare_input(self, frames_u8):
        """Input side of the loop body on the device (vae-kl.ipynb cell 2 + :211-212): uint8 grey levels [N,S,S] ->
        x = (frame/255 - 0.1307)/0.3081 as fp32 [N,1,S,S] and the int64 class targets y [N,S,S], one kernel."""
        self._require_cuda(frames_u8)
        f = frames_u8.contiguous()
        n = f.numel()
        x = torch.empty(f.shape[0], 1, *f.shape[1:], dtype=torch.float32, device=f.device)
        y = torch.empty(f.shape, dtype=torch.int64, device=f.device)
        check(lib.mmvae_prepare_input(_ptr(f), n, 255.0 * self.DATA_MEAN, 255.0 * self.DATA_STD, _ptr(x), _ptr(y), _stream()),
              "mmvae_prepare_input")
        return x, y

    def train_step(self, x, y, kl_weight=1.0, eps=None):
        """forward + loss + backward; returns the device tensor [loss, px_given_z, kl]"""
        self.forward(x, eps=eps, materialize=False)
        return self.loss_backward(y, kl_weight)

    @torch.no_grad()
    def decode(self, encoding):
        """decoder(encoding) (vae-kl.ipynb:162-166): [N,z,h,h] -> logits [N,classes,S,S]"""
        self._require_cuda(encoding)
        self._check_arena()
        encoding = encoding.to(torch.float32).contiguous()
        n = encoding.shape[0]
        desc, ws, info = self._workspace(n)
        recon = torch.empty(n, self.n_classes, self.image_size, self.image_size, dtype=torch.float32, device=encoding.device)
        check(lib.mmvae_decode(byref(desc), _ptr(encoding), _ptr(self._arena), c_void_p(0), c_void_p(0), _ptr(ws), ws.numel(),
                               _ptr(recon), _stream()), "mmvae_decode")
        self._state = None
        return recon
