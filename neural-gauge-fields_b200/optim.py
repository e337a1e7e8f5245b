"""``FusedAdam``: torch.optim.Adam's update (no weight decay, no amsgrad — what TriPlane/main.py:237 constructs with
``torch.optim.Adam(grad_vars, betas=(0.9, 0.99))``) as one kernel pass per parameter (``ngf_adam_step``).  It is a
``torch.optim.Optimizer``, so the reference's loop keeps working: ``zero_grad()``, ``step()``, per-group ``lr`` decay
(main.py:307-308), re-creation after ``up_sampling`` (main.py:355-357)."""
from __future__ import annotations

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam updates contiguous fp32 CUDA parameters only (there is no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad.contiguous().float()
                with torch.cuda.device(p.device):
                    _lib.check(lib.ngf_adam_step(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                                 p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                 int(st["step"]), int(torch.cuda.current_stream(p.device).cuda_stream)),
                               "ngf_adam_step")
                p._version  # noqa: B018  (the raw pointer write is invisible to autograd's version counter: bump it)
                p.add_(0)
        return loss
