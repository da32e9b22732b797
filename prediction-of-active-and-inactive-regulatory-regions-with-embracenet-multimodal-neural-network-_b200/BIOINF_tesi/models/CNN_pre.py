"""Mirror of BIOINF_tesi/models/CNN_pre.py:10-76 and CNN_pre_NoTrain.py:10-58 (parameter containers only)."""
import torch.nn as nn

from ...archspec import ArchSpec, size_out_convolution

MAXPOOL_KERNEL, MAXPOOL_STRIDE, INPUT_SIZE, IN_CHANNELS = 10, 2, 256, 4     # CNN_pre.py:18-22


def build_cnn_layers(channels, kernels, dropouts):
    """(Conv1d same-pad -> BatchNorm1d -> ReLU -> MaxPool1d(10,2) -> Dropout) x n: conv at 5i, BatchNorm at 5i+1."""
    layers, cin, size = [], IN_CHANNELS, INPUT_SIZE
    for co, k, p in zip(channels, kernels, dropouts):
        pad = int((k - 1) / 2)
        layers += [nn.Conv1d(cin, co, kernel_size=k, stride=1, padding=pad), nn.BatchNorm1d(co), nn.ReLU(),
                   nn.MaxPool1d(kernel_size=MAXPOOL_KERNEL, stride=MAXPOOL_STRIDE), nn.Dropout(p)]
        cin = co
        size = size_out_convolution(size_out_convolution(size, k, pad, 1), MAXPOOL_KERNEL, 0, MAXPOOL_STRIDE)
    return layers, cin * size


class CNN_pre(nn.Module):
    """One-hot-sequence docking feeder (parameter container; the engine runs the gather-sum first layer, the
    tcgen05 implicit-GEMM deeper layers and the fused BatchNorm/ReLU/MaxPool/Dropout kernels)."""

    def __init__(self, trial=None, device=None, channels=None, kernels=None, dropouts=None):
        super().__init__()
        self.trial, self.device = trial, device
        if channels is None:
            channels, kernels, dropouts = ArchSpec.suggest_cnn(trial, 'CNN_')
        self.channels, self.kernels, self.dropouts = list(channels), list(kernels), list(dropouts)
        layers, self.output_size = build_cnn_layers(self.channels, self.kernels, self.dropouts)
        self.CNN_model = nn.Sequential(*layers)

    def forward(self, x):
        raise RuntimeError('CNN_pre holds parameters for the B200 engine; call the owning model (no PyTorch fallback)')


class CNN_pre_NoTrain(CNN_pre):
    def __init__(self, model_params, device):
        n = int(model_params['n_layers'])
        super().__init__(None, device, channels=[model_params[f'out_channels_l{i}'] for i in range(n)],
                         kernels=[model_params[f'kernel_size_l{i}'] for i in range(n)],
                         dropouts=[model_params[f'dropout_l{i}'] for i in range(n)])
