"""Mirror of BIOINF_tesi/models/ConcatNetMultimodal.py:12-82 and ConcatNetMultimodal_NoTrain.py:11-90 — the
concatenation-fusion comparison model of every EmbraceNet notebook: post(cat(FFNN(x1), CNN(x2))).

Same constructor signature (note the argument order: in_features_FFNN BEFORE device), attribute names, forward signature
and state_dict keys; the arithmetic is engine kind EMB_KIND_CONCATNET: the two feeders' outputs are concatenated in the
engine's column order and the first post Linear runs with its weight columns permuted accordingly (the master weight
keeps the reference layout [ffnn_out | c * L + l])."""
import torch

from ...archspec import ArchSpec, CONCAT_UNITS
from ._base import EngineModule
from .CNN_pre import CNN_pre, CNN_pre_NoTrain
from .FFNN_pre import FFNN_pre, FFNN_pre_NoTrain
from .EmbraceNetMultimodal import _post_layers
from .utils.utils import get_single_model_params, output_size_from_model_params


class ConcatNetMultimodal(EngineModule):
    def __init__(self, trial, cell_line, task, in_features_FFNN, device, n_classes=2, args=None, embracenet_dropout=True,
                 precision=None, seed=0x5EED):
        super().__init__()
        if n_classes != 2:
            raise ValueError('the engine implements the reference\'s 2-class head')
        self.trial, self.cell_line, self.device, self.n_classes, self.args = trial, cell_line, device, n_classes, args
        self.FFNN = FFNN_pre(trial, in_features_FFNN, device=device)
        self.CNN = CNN_pre(trial, device=device)
        self.FFNN_pre_output_size, self.CNN_pre_output_size = self.FFNN.output_size, self.CNN.output_size
        units, drops = [], []
        for i in range(trial.suggest_int('CONCATNET_n_post_layers', 1, 3)):
            units.append(trial.suggest_categorical('CONCATNET_n_units_l{}'.format(i), CONCAT_UNITS[i]))
            drops.append(trial.suggest_categorical('CONCATNET_dropout_l{}'.format(i), [0.0, 0.2, 0.3, 0.5]))
        self.post = _post_layers(self.FFNN_pre_output_size + self.CNN_pre_output_size, units, drops, n_classes)
        spec = ArchSpec(kind='concatnet', in_features=int(in_features_FFNN), ffnn_units=self.FFNN.units, ffnn_dropout=self.FFNN.dropouts,
                        cnn_channels=self.CNN.channels, cnn_kernels=self.CNN.kernels, cnn_dropout=self.CNN.dropouts,
                        post_units=units, post_dropout=drops).validate()
        self._adopt(spec, device, precision, seed)

    def forward(self, x, draws=None):
        """model([x_FFNN, x_CNN]) -> logits [B, 2] (fp32); Dropout follows model.train() / model.eval()."""
        x_FFNN, x_CNN = x
        return self._run(x_FFNN, x_CNN, None, draws, modality_dropout=False)


class ConcatNetMultimodal_NoTrain(EngineModule):
    """Rebuilt from a checkpoint's `model_params` (ConcatNetMultimodal_NoTrain.py:11-90).  The reference computes a
    softmax there and drops it through a typo (`outuput`, :87), so its forward returns LOGITS: reproduced."""

    def __init__(self, cell_line, task, n_iter, in_features_FFNN, device, augmentation=False, n_classes=2, args=None,
                 precision=None, seed=0x5EED, model_params=None):
        super().__init__()
        self.cell_line, self.task, self.n_iter, self.device, self.n_classes = cell_line, task, n_iter, device, n_classes
        if model_params is None:
            aug = '_augmentation' if augmentation else ''
            saved = torch.load(f'{cell_line}_ConcatNetMultimodal{aug}_{task}_{n_iter}_test_.pt', map_location='cpu', weights_only=False)
            model_params = saved['model_params']
        single = get_single_model_params(model_params)
        self.FFNN = FFNN_pre_NoTrain(in_features_FFNN, single['FFNN'], device=device)
        self.CNN = CNN_pre_NoTrain(single['CNN'], device=device)
        for p in list(self.FFNN.parameters()) + list(self.CNN.parameters()):
            p.requires_grad = False
        self.FFNN_pre_output_size = single['FFNN'][f"n_units_l{single['FFNN']['n_layers'] - 1}"]
        self.CNN_pre_output_size = output_size_from_model_params(single['CNN'])
        n_post = model_params['CONCATNET_n_post_layers']
        units = [model_params[f'CONCATNET_n_units_l{i}'] for i in range(n_post)]
        drops = [model_params[f'CONCATNET_dropout_l{i}'] for i in range(n_post)]
        self.post = _post_layers(self.FFNN_pre_output_size + self.CNN_pre_output_size, units, drops, n_classes)
        self._adopt(ArchSpec.from_model_params(model_params, in_features_FFNN, kind='concatnet'), device, precision, seed)

    def forward(self, x, draws=None):
        x_FFNN, x_CNN = x
        return self._run(x_FFNN, x_CNN, None, draws, modality_dropout=False)
