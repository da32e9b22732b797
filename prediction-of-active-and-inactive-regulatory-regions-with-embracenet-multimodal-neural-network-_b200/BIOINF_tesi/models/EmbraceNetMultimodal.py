"""Mirror of BIOINF_tesi/models/EmbraceNetMultimodal.py:12-193 and EmbraceNetMultimodal_NoTrain.py:94-214.

Same constructor signature, attribute names, forward signature and state_dict keys as the reference; the arithmetic
runs in libembrace_sm100.so (docking GEMMs with the embracement select fused into the epilogue, ...)."""
import torch
import torch.nn as nn

from ...archspec import ArchSpec
from ._base import EngineModule
from .CNN_pre import CNN_pre, CNN_pre_NoTrain
from .FFNN_pre import FFNN_pre, FFNN_pre_NoTrain
from .utils.utils import get_single_model_params, output_size_from_model_params


class EmbraceNet(nn.Module):
    """Parameter container of EmbraceNetMultimodal.py:12-32: docking_i = nn.Linear(input_size_i, embracement_size)."""

    def __init__(self, device, input_size_list, embracement_size=256, bypass_docking=False):
        super().__init__()
        if bypass_docking:
            raise NotImplementedError('bypass_docking is never used by the reference (EmbraceNetMultimodal.py:127-129)')
        self.device, self.input_size_list, self.embracement_size = device, input_size_list, embracement_size
        self.bypass_docking = bypass_docking
        for i, input_size in enumerate(input_size_list):
            setattr(self, 'docking_%d' % i, nn.Linear(input_size, embracement_size))

    def forward(self, *a, **k):
        raise RuntimeError('EmbraceNet holds parameters for the B200 engine; call the owning model')


def _post_layers(in_features, units, dropouts, n_classes):
    layers = []
    for u, p in zip(units, dropouts):
        layers += [nn.Linear(in_features, u), nn.ReLU(), nn.Dropout(p)]
        in_features = u
    layers.append(nn.Linear(in_features, n_classes))
    return nn.Sequential(*layers)


class EmbraceNetMultimodal(EngineModule):
    def __init__(self, trial, cell_line, task, device, in_features_FFNN, n_classes=2, args=None, embracenet_dropout=True,
                 precision=None, seed=0x5EED):
        super().__init__()
        if n_classes != 2:
            raise ValueError('the engine implements the reference\'s 2-class head')
        self.trial, self.cell_line, self.task, self.device = trial, cell_line, task, device
        self.n_classes, self.embracenet_dropout, self.args = n_classes, embracenet_dropout, args
        # same suggestion order as the reference: FFNN, CNN, embracement size, post layers, selection probability
        self.FFNN = FFNN_pre(trial, in_features_FFNN, device=device)
        self.CNN = CNN_pre(trial, device=device)
        self.FFNN_pre_output_size, self.CNN_pre_output_size = self.FFNN.output_size, self.CNN.output_size
        C = trial.suggest_categorical('EMBRACENET_embracement_size', [512, 768, 1024])
        self.embracenet = EmbraceNet(device, [self.FFNN_pre_output_size, self.CNN_pre_output_size], embracement_size=C)
        n_post = trial.suggest_int('n_post_layers', 0, 2)
        units, drops = [], []
        for i in range(n_post):
            units.append(trial.suggest_categorical(f'EMBRACENET_n_units_l{i}', [32, 64, 128, 256, 512] if i == 0 else [16, 32, 64, 128, 256]))
            drops.append(trial.suggest_categorical(f'EMBRACENET_dropout_l{i}', [0.0, 0.2, 0.3, 0.5]))
        self.post = _post_layers(C, units, drops, n_classes)
        p = trial.suggest_float('selection_probabilities_FFNN', 0.0, 1.0)
        self.selection_probabilities = torch.tensor([p, 1.0 - p])       # plain attribute, not a buffer (quirk)
        spec = ArchSpec(kind='embracenet', in_features=int(in_features_FFNN), ffnn_units=self.FFNN.units, ffnn_dropout=self.FFNN.dropouts,
                        cnn_channels=self.CNN.channels, cnn_kernels=self.CNN.kernels, cnn_dropout=self.CNN.dropouts, embracement_size=C,
                        post_units=units, post_dropout=drops, p_ffnn=float(p), embracenet_dropout=True).validate()
        self._adopt(spec, device, precision, seed)

    def forward(self, x, availabilities=None, selection_probabilities=None, is_training=False, embracenet_dropout=True, draws=None):
        """model([x_FFNN, x_CNN], ...) -> logits [B,2] (fp32).  As in the reference the `selection_probabilities`
        argument is ignored (EmbraceNetMultimodal.py:184-187); `draws` (optional) replays explicit random draws."""
        x_FFNN, x_CNN = x
        return self._run(x_FFNN, x_CNN, availabilities, draws, modality_dropout=bool(is_training and embracenet_dropout))


class EmbraceNetMultimodal_NoTrain(EngineModule):
    """Rebuilt from a checkpoint's `model_params` (EmbraceNetMultimodal_NoTrain.py:118-177); forward returns the
    flattened softmax, so `model_([x1, x2])[1]` is P(class 1) of a single sample as in visual.py:290-293.
    `predict_proba(x1, x2)` is the batched form of that loop."""

    def __init__(self, cell_line, task, n_iter, in_features_FFNN, device, augmentation=False, n_classes=2, args=None,
                 embracenet_dropout=True, precision=None, seed=0x5EED, model_params=None):
        super().__init__()
        self.cell_line, self.task, self.n_iter, self.device = cell_line, task, n_iter, device
        self.n_classes, self.embracenet_dropout, self.args = n_classes, embracenet_dropout, args
        self.softmax_layer = torch.nn.Softmax(dim=None)
        if model_params is None:
            aug = '_augmentation' if augmentation else ''
            saved = torch.load(f'{cell_line}_EmbraceNetMultimodal{aug}_{task}_{n_iter}_test_.pt', map_location='cpu', weights_only=False)
            model_params = saved['model_params']
        single = get_single_model_params(model_params)
        self.FFNN = FFNN_pre_NoTrain(in_features_FFNN, single['FFNN'], device=device)
        self.CNN = CNN_pre_NoTrain(single['CNN'], device=device)
        for p in list(self.FFNN.parameters()) + list(self.CNN.parameters()):
            p.requires_grad = False
        self.FFNN_pre_output_size = single['FFNN'][f"n_units_l{single['FFNN']['n_layers'] - 1}"]
        self.CNN_pre_output_size = output_size_from_model_params(single['CNN'])
        C = model_params['EMBRACENET_embracement_size']
        self.embracenet = EmbraceNet(device, [self.FFNN_pre_output_size, self.CNN_pre_output_size], embracement_size=C)
        n_post = model_params['n_post_layers']
        units = [model_params[f'EMBRACENET_n_units_l{i}'] for i in range(n_post)]
        drops = [model_params[f'EMBRACENET_dropout_l{i}'] for i in range(n_post)]
        self.post = _post_layers(C, units, drops, n_classes)
        p = model_params['selection_probabilities_FFNN']
        self.selection_probabilities = torch.tensor([p, 1.0 - p])
        self._adopt(ArchSpec.from_model_params(model_params, in_features_FFNN), device, precision, seed)

    def forward(self, x, availabilities=None, selection_probabilities=None, is_training=False, embracenet_dropout=True, draws=None):
        x_FFNN, x_CNN = x
        logits = self._run(x_FFNN, x_CNN, availabilities, draws, modality_dropout=bool(is_training and embracenet_dropout))
        return torch.softmax(logits, dim=1).reshape(-1)

    def predict_proba(self, x_FFNN, x_CNN, availabilities=None, batch_size=65536, draws=None):
        """P(class 1) for every row, batched (replaces the batch-1 Python loop of visual.py:290-293)."""
        return self.predict_scores(x_FFNN, x_CNN, availabilities, batch_size, draws, column='prob')
