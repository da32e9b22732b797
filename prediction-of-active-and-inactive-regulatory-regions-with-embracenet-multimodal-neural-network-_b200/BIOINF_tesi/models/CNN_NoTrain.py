"""Mirror of BIOINF_tesi/models/CNN_NoTrain.py:11-81: the single-modality CNN rebuilt from a checkpoint's `model_params`
(conv stack + Linear(->1000) -> Linear(->64) -> Linear(->classes), no activations between them); forward returns the
flattened softmax."""
import torch
import torch.nn as nn

from ...archspec import ArchSpec
from ._base import EngineModule
from .CNN_pre import build_cnn_layers


class CNN_NoTrain(EngineModule):
    def __init__(self, cell_line, task, n_iter, device, classes=2, precision=None, seed=0x5EED, model_params=None):
        super().__init__()
        self.cell_line, self.task, self.n_iter, self.device, self.classes = cell_line, task, n_iter, device, classes
        self.softmax_layer = torch.nn.Softmax(dim=None)
        if model_params is None:
            saved = torch.load(f'{cell_line}_CNN_{task}_{n_iter}_test_.pt', map_location='cpu', weights_only=False)
            model_params = saved['model_params']
        spec = ArchSpec.from_model_params(model_params, 0, kind='cnn')
        layers, out = build_cnn_layers(spec.cnn_channels, spec.cnn_kernels, spec.cnn_dropout)
        self.CNN_model = nn.Sequential(*layers)
        self.last_layer1 = nn.Linear(out, 1000)
        self.last_layer2 = nn.Linear(1000, 64)
        self.last_output = nn.Linear(64, classes)
        self._adopt(spec, device, precision, seed)

    def forward(self, x, draws=None):
        logits = self._run(None, x, None, draws, modality_dropout=False)
        return torch.softmax(logits, dim=1).reshape(-1)

    def predict_proba(self, x, batch_size=65536):
        """P(class 1) for every row (batched form of `[model_(X_2.loc[i])[1] for i ...]`, visual.py:286-287)."""
        return self.predict_scores(None, x, None, batch_size, None, column='prob')
