"""Mirror of BIOINF_tesi/models/CNN_net.py:10-83 (single-modality CNN baseline; three Linear layers without
activations after the conv stack, CNN_net.py:71-81)."""
import torch.nn as nn

from ...archspec import ArchSpec
from ._base import EngineModule
from .CNN_pre import build_cnn_layers


class CNN(EngineModule):
    def __init__(self, trial, device, classes=2, precision=None, seed=0x5EED):
        super().__init__()
        self.trial, self.classes, self.device = trial, classes, device
        ch, ks, drops = ArchSpec.suggest_cnn(trial, '')
        layers, out = build_cnn_layers(ch, ks, drops)
        self.CNN_model = nn.Sequential(*layers)
        self.last_layer1 = nn.Linear(out, 1000)
        self.last_layer2 = nn.Linear(1000, 64)
        self.last_output = nn.Linear(64, classes)
        self._adopt(ArchSpec(kind='cnn', cnn_channels=ch, cnn_kernels=ks, cnn_dropout=drops).validate(), device, precision, seed)

    def forward(self, x, draws=None):
        return self._run(None, x, None, draws, modality_dropout=False)
