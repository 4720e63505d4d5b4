"""Fused Adam over the flat parameter arena: the step right after the hot path (`optim.Adam(list(model.parameters()))`
main.py:468, `optimizer.step()` main.py:399).

torch.optim.Adam walks ~93 parameter tensors (multi-tensor apply: several launches, per-tensor state); here every
nn.Parameter of a mmvae_b200 module is a view of ONE fp32 arena and autograd leaves every `.grad` as a view of ONE flat
gradient buffer, so the whole update is a single elementwise kernel (`mmvae_adam_step`): m, v, bias correction, the
parameter write -- and the 1/world_size of a data-parallel sum -- in one pass over 4 x 8.45 MB.

Semantics are torch.optim.Adam's defaults (amsgrad=False, maximize=False, L2 weight decay added to the gradient):
`tests/test_gpu_adam.py` compares ten steps against torch.optim.Adam.  There is no CPU path.
"""
import ctypes

import torch

from ._lib import check, lib


class FusedAdam:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        """`model`: a mmvae_b200.VAE / NotebookVAE on a CUDA device (anything with `flat_parameters` whose parameters
        are views of it).  `grad_scale` multiplies the gradient before the update (1/world_size if the exchange summed)."""
        flat = model.flat_parameters
        if not flat.is_cuda:
            raise ValueError("FusedAdam needs the module on a CUDA device (there is no CPU path)")
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.grad_scale = lr, betas, eps, weight_decay, grad_scale
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = 0
        # the step count also lives on the device: a captured CUDA graph computes the bias corrections from it
        self._step_dev = torch.zeros((), dtype=torch.int64, device=flat.device)
        self._scratch_grad = None

    # -- gradients ---------------------------------------------------------------------------------------------
    def _flat_grad(self):
        """The flat gradient buffer the parameters' `.grad` tensors are views of; gathered into a scratch buffer when
        they are not (e.g. a gradient was accumulated over two backward calls into fresh tensors)."""
        m = self.model
        params = list(m.parameters())
        g = getattr(m, "last_flat_grad", None)
        flat = m.flat_parameters
        if g is not None and g.numel() == flat.numel():
            base_p, base_g = flat.data_ptr(), g.data_ptr()
            if all(p.grad is not None and p.grad.data_ptr() - base_g == p.data_ptr() - base_p for p in params):
                return g
        if any(p.grad is None for p in params):
            raise RuntimeError("FusedAdam.step(): a parameter has no gradient (call backward() first)")
        if self._scratch_grad is None:
            self._scratch_grad = torch.empty_like(flat)
        base_p = flat.data_ptr()
        for p in params:
            off = (p.data_ptr() - base_p) // 4
            self._scratch_grad[off:off + p.numel()].copy_(p.grad.reshape(-1))
        return self._scratch_grad

    # -- optimizer API -----------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self, flat_grad=None):
        m = self.model
        if hasattr(m, "_check_arena"):
            m._check_arena()
        flat = m.flat_parameters
        if flat.data_ptr() != self.exp_avg.data_ptr() and flat.numel() != self.exp_avg.numel():
            raise RuntimeError("the parameter arena changed size since FusedAdam was built")
        g = flat_grad if flat_grad is not None else self._flat_grad()
        self.step_count += 1
        self._step_dev += 1
        P = ctypes.c_void_p
        with torch.cuda.device(flat.device):
            check(lib.mmvae_adam_step(flat.numel(), P(flat.data_ptr()), P(g.data_ptr()), P(self.exp_avg.data_ptr()),
                                      P(self.exp_avg_sq.data_ptr()), self.lr, self.betas[0], self.betas[1], self.eps,
                                      self.weight_decay, self.step_count, P(self._step_dev.data_ptr()), self.grad_scale,
                                      P(torch.cuda.current_stream(flat.device).cuda_stream)), "mmvae_adam_step")

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self._step_dev.fill_(self.step_count)
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["weight_decay"]
