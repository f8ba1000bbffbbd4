"""FusedAdam -- torch.optim.Adam semantics (trainer/trainer.py:45-49: Adam(lr, weight_decay), no
amsgrad) as ONE pass over (p, g, m, v) per parameter with the native kernel; when the parameters are
adjacent views of a flat table and so are their gradients, one launch covers them all."""
from __future__ import annotations

import torch

from ._lib import check, lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group['betas']
            todo = []
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError('FusedAdam: CUDA float32 parameters only (no CPU path)')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                todo.append((p, g, st))
            for p, g, st in todo:
                if not p.is_contiguous():
                    raise RuntimeError('FusedAdam: parameters must be contiguous')
                with torch.cuda.device(p.device):
                    check(lib.ssl_adam_step(p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                                            p.numel(), st['step'], group['lr'], b1, b2, group['eps'], group['weight_decay'],
                                            torch.cuda.current_stream(p.device).cuda_stream), 'ssl_adam_step')
        return loss
