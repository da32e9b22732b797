"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, plans shapes like the reference, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import embrace_b200
from embrace_b200 import _native as N
from embrace_b200.archspec import ArchSpec
from oracle import embracenet_oracle as O
from tests.golden.cases import CASES, ARCH_S, ARCH_L, ARCH_M, ARCH_W
from tests.test_gpu_parity import to_archspec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    embrace_b200.build()
    return embrace_b200.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'embrace_b200.h')).read()
    declared = set(re.findall(r'\b(emb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.emb_abi_version() == 2


@pytest.mark.parametrize('spec', [ARCH_S, ARCH_M, ARCH_L, ARCH_W] + [c['spec'] for c in CASES.values()])
def test_param_table_matches_reference_state_dict(lib, spec):
    h = C.c_void_p()
    cs = to_archspec(spec).to_c()
    N.check(lib.emb_create(C.byref(cs), 16, 0, C.byref(h)))
    shapes = {k: v for k, v in O.param_shapes(spec).items() if not k.endswith('num_batches_tracked')}
    info = N.EmbParamInfo()
    got = {}
    for i in range(lib.emb_num_tensors(h)):
        N.check(lib.emb_param_info(h, i, C.byref(info)))
        got[info.name.decode()] = tuple(info.shape[:info.ndim])
        assert info.offset % 4 == 0
    assert list(got) == list(shapes)          # same keys, same order as the reference nn.Sequential
    assert got == {k: tuple(v) for k, v in shapes.items()}
    a = to_archspec(spec)
    if spec.get('kind', 'embracenet') != 'cnn':
        assert lib.emb_output_size(h, 0) == a.ffnn_output_size
    if spec.get('kind', 'embracenet') != 'ffnn':
        assert lib.emb_output_size(h, 1) == a.cnn_output_size
    lib.emb_destroy(h)


def test_known_sizes():
    assert to_archspec(ARCH_S).cnn_output_size == 1600
    assert to_archspec(ARCH_M).cnn_output_size == 1024
    assert to_archspec(ARCH_L).cnn_output_size == 4096
    assert to_archspec(ARCH_W).cnn_output_size == 7936


def test_bad_specs_are_rejected(lib):
    h = C.c_void_p()
    s = to_archspec(ARCH_S)
    s.cnn_kernels = [4, 5, 5]
    with pytest.raises(ValueError):
        s.validate()
    cs = to_archspec(ARCH_S).to_c()
    cs.cnn_kernels[0] = 4
    assert lib.emb_create(C.byref(cs), 16, 0, C.byref(h)) == -1
    assert b'odd' in lib.emb_last_error()
    cs = to_archspec(ARCH_S).to_c()
    assert lib.emb_create(C.byref(cs), 0, 0, C.byref(h)) == -1
    assert lib.emb_create(C.byref(cs), 8, 7, C.byref(h)) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.emb_device_count() == 0
    with pytest.raises(N.EmbError):
        embrace_b200.Engine(to_archspec(ARCH_S), 8)
    h = C.c_void_p()
    cs = to_archspec(ARCH_S).to_c()
    N.check(lib.emb_create(C.byref(cs), 8, 0, C.byref(h)))
    assert lib.emb_bind(h, None, None, None, None, None, None, 0) == -2      # EMB_E_NO_DEVICE
    assert lib.emb_forward_train(h, None, None, None, 4, None, None, None) == -4  # EMB_E_STATE: nothing bound
    lib.emb_destroy(h)


class _Trial:
    def __init__(self, params):
        self.params, self.asked = params, []

    def suggest_int(self, name, lo, hi):
        self.asked.append(name)
        return self.params[name]

    def suggest_categorical(self, name, choices):
        self.asked.append(name)
        assert self.params[name] in choices, (name, choices)
        return self.params[name]

    def suggest_float(self, name, lo, hi):
        self.asked.append(name)
        return self.params[name]


def test_archspec_from_trial_asks_like_the_reference():
    a = to_archspec(ARCH_S)
    mp = a.to_model_params()
    t = _Trial(mp)
    b = ArchSpec.from_trial(t, 48)
    assert b == a
    # the reference's suggestion order: FFNN (layers, then units/dropout per layer), CNN, embracement, post, p
    assert t.asked[0] == 'FFNN_n_layers' and t.asked[1:3] == ['FFNN_n_units_l0', 'FFNN_dropout_l0']
    assert t.asked[5] == 'CNN_n_layers' and t.asked[6:9] == ['CNN_out_channels_l0', 'CNN_kernel_size_l0', 'CNN_dropout_l0']
    assert t.asked[-1] == 'selection_probabilities_FFNN'
    assert ArchSpec.from_model_params(mp, 48) == a
