"""Mirror of BIOINF_tesi/models/FF_net.py:8-50 (single-modality FFNN baseline)."""
import torch.nn as nn

from ...archspec import ArchSpec
from ._base import EngineModule
from .FFNN_pre import build_ffnn_layers


class FFNN(EngineModule):
    def __init__(self, trial, in_features, device, classes=2, precision=None, seed=0x5EED):
        super().__init__()
        self.trial, self.classes, self.device = trial, classes, device
        units, drops = ArchSpec.suggest_ffnn(trial, '')
        layers, last = build_ffnn_layers(in_features, units, drops)
        layers.append(nn.Linear(last, classes))
        self.model = nn.Sequential(*layers)
        self._adopt(ArchSpec(kind='ffnn', in_features=int(in_features), ffnn_units=units, ffnn_dropout=drops).validate(), device, precision, seed)

    def forward(self, x, draws=None):
        return self._run(x, None, None, draws, modality_dropout=False)
