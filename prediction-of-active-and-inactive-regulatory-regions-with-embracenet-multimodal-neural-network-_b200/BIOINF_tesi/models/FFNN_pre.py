"""Mirror of BIOINF_tesi/models/FFNN_pre.py:8-49 and FFNN_pre_NoTrain.py:8-35 (parameter containers only)."""
import torch.nn as nn


def build_ffnn_layers(in_features, units, dropouts):
    """(Linear -> ReLU -> Dropout) x n, exactly the reference's nn.Sequential indices (weights at 0, 3, 6, 9)."""
    layers = []
    for u, p in zip(units, dropouts):
        layers += [nn.Linear(in_features, u), nn.ReLU(), nn.Dropout(p)]
        in_features = u
    return layers, in_features


class FFNN_pre(nn.Module):
    """Epigenomic docking feeder.  Inside EmbraceNetMultimodal it only holds parameters; its arithmetic runs in
    the engine's Linear kernels (bias/ReLU/Dropout fused into the GEMM epilogue)."""

    def __init__(self, trial=None, in_features=None, device=None, classes=2, units=None, dropouts=None):
        super().__init__()
        self.trial, self.classes, self.device = trial, classes, device
        if units is None:
            from ...archspec import ArchSpec
            units, dropouts = ArchSpec.suggest_ffnn(trial, 'FFNN_')
        self.units, self.dropouts = list(units), list(dropouts)
        layers, self.output_size = build_ffnn_layers(in_features, self.units, self.dropouts)
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        raise RuntimeError('FFNN_pre holds parameters for the B200 engine; call the owning model (no PyTorch fallback)')


class FFNN_pre_NoTrain(FFNN_pre):
    def __init__(self, in_features, model_params, device):
        n = int(model_params['n_layers'])
        super().__init__(None, in_features, device, units=[model_params[f'n_units_l{i}'] for i in range(n)],
                         dropouts=[model_params[f'dropout_l{i}'] for i in range(n)])
