"""Mirror of BIOINF_tesi/models/FFNN_NoTrain.py:8-51: the single-modality FFNN rebuilt from a checkpoint's `model_params`
for the predict loop of visual.py:263-295; forward returns the flattened softmax."""
import torch
import torch.nn as nn

from ...archspec import ArchSpec
from ._base import EngineModule
from .FFNN_pre import build_ffnn_layers


class FFNN_NoTrain(EngineModule):
    def __init__(self, cell_line, task, n_iter, in_features, device, classes=2, precision=None, seed=0x5EED, model_params=None):
        super().__init__()
        self.cell_line, self.task, self.n_iter, self.device, self.classes = cell_line, task, n_iter, device, classes
        self.softmax_layer = torch.nn.Softmax(dim=None)
        if model_params is None:
            saved = torch.load(f'{cell_line}_FFNN_{task}_{n_iter}_test_.pt', map_location='cpu', weights_only=False)
            model_params = saved['model_params']
        spec = ArchSpec.from_model_params(model_params, in_features, kind='ffnn')
        layers, last = build_ffnn_layers(in_features, spec.ffnn_units, spec.ffnn_dropout)
        layers.append(nn.Linear(last, classes))
        self.model = nn.Sequential(*layers)
        self._adopt(spec, device, precision, seed)

    def forward(self, x, draws=None):
        logits = self._run(x, None, None, draws, modality_dropout=False)
        return torch.softmax(logits, dim=1).reshape(-1)

    def predict_proba(self, x, batch_size=65536):
        """P(class 1) for every row (batched form of `[model_(X_1.loc[i])[1] for i ...]`, visual.py:284-285)."""
        return self.predict_scores(x, None, None, batch_size, None, column='prob')
