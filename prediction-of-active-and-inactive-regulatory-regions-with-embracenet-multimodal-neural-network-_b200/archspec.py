"""The architecture an Optuna trial / a checkpoint's `model_params` dict selects, as a frozen value.

Mirrors the `trial.suggest_*` calls of the reference constructors, in the reference's order and with the
reference's parameter names, so that a duck-typed trial sees exactly the same sequence of suggestions:
  FFNN_pre.py:19-39, CNN_pre.py:24-51, EmbraceNetMultimodal.py:123-157 (prefixed names)
  FF_net.py:19-39, CNN_net.py:26-52 (un-prefixed names)
and the `model_params` readers of FFNN_pre_NoTrain.py:15-28, CNN_pre_NoTrain.py:24-46,
EmbraceNetMultimodal_NoTrain.py:123-177.
"""
from dataclasses import dataclass, field
from typing import List

from . import _native as N

FFNN_UNITS = [[32, 64, 128, 256], [16, 32, 64, 128], [4, 16, 32, 64], [4, 16, 32]]
CNN_CHANNELS = [[16, 32, 64], [32, 64, 96], [64, 96, 128, 256], [128, 256, 512]]
CNN_KERNELS = [5, 11, 15]
POST_UNITS = [[32, 64, 128, 256, 512], [16, 32, 64, 128, 256]]
CONCAT_UNITS = [[512, 768, 1024], [32, 64, 128, 256, 512], [16, 32, 64, 128, 256]]     # ConcatNetMultimodal.py:46-52
SEQ_LEN, POOL_K, POOL_S = 256, 10, 2


def size_out_convolution(input_size, kernel, padding, stride):
    """utils.py:143-153."""
    return int(((input_size + 2 * padding - kernel) / stride) + 1)


@dataclass
class ArchSpec:
    kind: str = 'embracenet'
    in_features: int = 0
    ffnn_units: List[int] = field(default_factory=list)
    ffnn_dropout: List[float] = field(default_factory=list)
    cnn_channels: List[int] = field(default_factory=list)
    cnn_kernels: List[int] = field(default_factory=list)
    cnn_dropout: List[float] = field(default_factory=list)
    embracement_size: int = 0
    post_units: List[int] = field(default_factory=list)
    post_dropout: List[float] = field(default_factory=list)
    p_ffnn: float = 0.5
    embracenet_dropout: bool = True

    # ---- constructors ---------------------------------------------------------------------
    @staticmethod
    def suggest_ffnn(trial, prefix):
        n = trial.suggest_int(f'{prefix}n_layers', 1, 4)
        units, drops = [], []
        for i in range(n):
            units.append(trial.suggest_categorical(f'{prefix}n_units_l{i}', FFNN_UNITS[i]))
            drops.append(trial.suggest_categorical(f'{prefix}dropout_l{i}', [0.0, 0.2, 0.3, 0.4] if i < 2 else [0.0, 0.4, 0.5]))
        return units, drops

    @staticmethod
    def suggest_cnn(trial, prefix):
        n = trial.suggest_int(f'{prefix}n_layers', 1, 4)
        ch, ks, drops = [], [], []
        for i in range(n):
            ch.append(trial.suggest_categorical(f'{prefix}out_channels_l{i}', CNN_CHANNELS[i]))
            ks.append(trial.suggest_categorical(f'{prefix}kernel_size_l{i}', CNN_KERNELS))
            drops.append(trial.suggest_categorical(f'{prefix}dropout_l{i}', [0, 0.2, 0.3, 0.4] if i < 1 else [0, 0.4, 0.5]))
        return ch, ks, drops

    @classmethod
    def from_trial(cls, trial, in_features_FFNN, kind='embracenet', embracenet_dropout=True):
        s = cls(kind=kind, in_features=int(in_features_FFNN or 0), embracenet_dropout=embracenet_dropout)
        if kind == 'embracenet':
            s.ffnn_units, s.ffnn_dropout = cls.suggest_ffnn(trial, 'FFNN_')
            s.cnn_channels, s.cnn_kernels, s.cnn_dropout = cls.suggest_cnn(trial, 'CNN_')
            s.embracement_size = trial.suggest_categorical('EMBRACENET_embracement_size', [512, 768, 1024])
            n_post = trial.suggest_int('n_post_layers', 0, 2)
            for i in range(n_post):
                s.post_units.append(trial.suggest_categorical(f'EMBRACENET_n_units_l{i}', POST_UNITS[i]))
                s.post_dropout.append(trial.suggest_categorical(f'EMBRACENET_dropout_l{i}', [0.0, 0.2, 0.3, 0.5]))
            s.p_ffnn = trial.suggest_float('selection_probabilities_FFNN', 0.0, 1.0)
        elif kind == 'concatnet':          # ConcatNetMultimodal.py:33-58
            s.ffnn_units, s.ffnn_dropout = cls.suggest_ffnn(trial, 'FFNN_')
            s.cnn_channels, s.cnn_kernels, s.cnn_dropout = cls.suggest_cnn(trial, 'CNN_')
            for i in range(trial.suggest_int('CONCATNET_n_post_layers', 1, 3)):
                s.post_units.append(trial.suggest_categorical(f'CONCATNET_n_units_l{i}', CONCAT_UNITS[i]))
                s.post_dropout.append(trial.suggest_categorical(f'CONCATNET_dropout_l{i}', [0.0, 0.2, 0.3, 0.5]))
        elif kind == 'ffnn':
            s.ffnn_units, s.ffnn_dropout = cls.suggest_ffnn(trial, '')
        elif kind == 'cnn':
            s.cnn_channels, s.cnn_kernels, s.cnn_dropout = cls.suggest_cnn(trial, '')
        else:
            raise ValueError(kind)
        return s.validate()

    @classmethod
    def from_model_params(cls, model_params, in_features_FFNN, kind='embracenet', embracenet_dropout=True):
        """model_params: the trial-parameter dict stored in a reference checkpoint."""
        mp = model_params
        s = cls(kind=kind, in_features=int(in_features_FFNN or 0), embracenet_dropout=embracenet_dropout)
        pf, pc = ('FFNN_', 'CNN_') if kind in ('embracenet', 'concatnet') else ('', '')
        if kind in ('embracenet', 'concatnet', 'ffnn'):
            for i in range(int(mp[f'{pf}n_layers'])):
                s.ffnn_units.append(int(mp[f'{pf}n_units_l{i}']))
                s.ffnn_dropout.append(float(mp[f'{pf}dropout_l{i}']))
        if kind in ('embracenet', 'concatnet', 'cnn'):
            for i in range(int(mp[f'{pc}n_layers'])):
                s.cnn_channels.append(int(mp[f'{pc}out_channels_l{i}']))
                s.cnn_kernels.append(int(mp[f'{pc}kernel_size_l{i}']))
                s.cnn_dropout.append(float(mp[f'{pc}dropout_l{i}']))
        if kind == 'embracenet':
            s.embracement_size = int(mp['EMBRACENET_embracement_size'])
            for i in range(int(mp['n_post_layers'])):
                s.post_units.append(int(mp[f'EMBRACENET_n_units_l{i}']))
                s.post_dropout.append(float(mp[f'EMBRACENET_dropout_l{i}']))
            s.p_ffnn = float(mp['selection_probabilities_FFNN'])
        if kind == 'concatnet':
            for i in range(int(mp['CONCATNET_n_post_layers'])):
                s.post_units.append(int(mp[f'CONCATNET_n_units_l{i}']))
                s.post_dropout.append(float(mp[f'CONCATNET_dropout_l{i}']))
        return s.validate()

    def to_model_params(self):
        pf, pc = ('FFNN_', 'CNN_') if self.kind in ('embracenet', 'concatnet') else ('', '')
        mp = {}
        if self.kind in ('embracenet', 'concatnet', 'ffnn'):
            mp[f'{pf}n_layers'] = len(self.ffnn_units)
            for i, (u, p) in enumerate(zip(self.ffnn_units, self.ffnn_dropout)):
                mp[f'{pf}n_units_l{i}'], mp[f'{pf}dropout_l{i}'] = u, p
        if self.kind in ('embracenet', 'concatnet', 'cnn'):
            mp[f'{pc}n_layers'] = len(self.cnn_channels)
            for i, (c, k, p) in enumerate(zip(self.cnn_channels, self.cnn_kernels, self.cnn_dropout)):
                mp[f'{pc}out_channels_l{i}'], mp[f'{pc}kernel_size_l{i}'], mp[f'{pc}dropout_l{i}'] = c, k, p
        if self.kind == 'embracenet':
            mp['EMBRACENET_embracement_size'] = self.embracement_size
            mp['n_post_layers'] = len(self.post_units)
            for i, (u, p) in enumerate(zip(self.post_units, self.post_dropout)):
                mp[f'EMBRACENET_n_units_l{i}'], mp[f'EMBRACENET_dropout_l{i}'] = u, p
            mp['selection_probabilities_FFNN'] = self.p_ffnn
        if self.kind == 'concatnet':
            mp['CONCATNET_n_post_layers'] = len(self.post_units)
            for i, (u, p) in enumerate(zip(self.post_units, self.post_dropout)):
                mp[f'CONCATNET_n_units_l{i}'], mp[f'CONCATNET_dropout_l{i}'] = u, p
        return mp

    # ---- derived shapes --------------------------------------------------------------------
    def validate(self):
        if self.kind not in N.KIND:
            raise ValueError(f'unknown model kind {self.kind!r}')
        if len(self.ffnn_units) > 4 or len(self.cnn_channels) > 4 or len(self.post_units) > (3 if self.kind == 'concatnet' else 2):
            raise ValueError('too many layers (FFNN<=4, CNN<=4, post<=2; ConcatNet post<=3)')
        for k in self.cnn_kernels:
            if k % 2 == 0:
                raise ValueError('only odd kernel sizes ("same" padding) are supported')
        return self

    @property
    def ffnn_output_size(self):
        return self.ffnn_units[-1] if self.ffnn_units else 0

    def cnn_lengths(self):
        out, L = [], SEQ_LEN
        for k in self.cnn_kernels:
            p = int((k - 1) / 2)
            Lc = size_out_convolution(L, k, p, 1)
            Lp = size_out_convolution(Lc, POOL_K, 0, POOL_S)
            out.append((Lc, Lp))
            L = Lp
        return out

    @property
    def cnn_output_size(self):
        return self.cnn_channels[-1] * self.cnn_lengths()[-1][1] if self.cnn_channels else 0

    def to_c(self):
        c = N.EmbArchSpec()
        c.kind = N.KIND[self.kind]
        c.in_features = self.in_features
        c.n_ffnn = len(self.ffnn_units)
        for i, (u, p) in enumerate(zip(self.ffnn_units, self.ffnn_dropout)):
            c.ffnn_units[i], c.ffnn_dropout[i] = u, p
        c.n_cnn = len(self.cnn_channels)
        for i, (ch, k, p) in enumerate(zip(self.cnn_channels, self.cnn_kernels, self.cnn_dropout)):
            c.cnn_channels[i], c.cnn_kernels[i], c.cnn_dropout[i] = ch, k, p
        c.embracement_size = self.embracement_size
        c.n_post = len(self.post_units)
        for i, (u, p) in enumerate(zip(self.post_units, self.post_dropout)):
            c.post_units[i], c.post_dropout[i] = u, p
        c.p_ffnn = self.p_ffnn
        c.embracenet_dropout = 1 if self.embracenet_dropout else 0
        return c
