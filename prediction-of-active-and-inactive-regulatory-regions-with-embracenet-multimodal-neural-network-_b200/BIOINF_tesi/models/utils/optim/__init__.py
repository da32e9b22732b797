"""Mirror of BIOINF_tesi/models/utils/optim/__init__.py:1-4.  timm is not a dependency: `Nadam` is torch's NAdam with
timm's schedule_decay default (same published rule); fit()/fit_multimodal() recognise these classes and run
the engine's fused optimizer kernel with their hyper-parameters."""
import torch
from torch.optim import Adam, RMSprop


class Nadam(torch.optim.NAdam):
    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, schedule_decay=4e-3):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, momentum_decay=schedule_decay)


__all__ = ['Adam', 'RMSprop', 'Nadam']
