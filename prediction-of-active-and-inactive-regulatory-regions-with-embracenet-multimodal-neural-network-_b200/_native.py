"""ctypes binding of libembrace_sm100.so (include/embrace_b200.h) and its in-tree build.

The library is the product: there is no Python/PyTorch fallback for any compute entry.  If the
shared object is missing, or no sm_100 device is visible, the host classes raise.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libembrace_sm100.so')
INCLUDE = os.path.join(REPO, 'include')

EMB_MAX_FFNN, EMB_MAX_CNN, EMB_MAX_POST = 4, 4, 3
KIND = {'embracenet': 0, 'ffnn': 1, 'cnn': 2, 'concatnet': 3}
PREC = {'fp32': 0, 'bf16': 1}
OPT = {'adam': 0, 'adamw': 1, 'nadam': 2, 'rmsprop': 3}

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared',
              '-Xcompiler', '-fPIC', '-I', INCLUDE]


class EmbArchSpec(C.Structure):
    _fields_ = [('kind', C.c_int32), ('in_features', C.c_int32),
                ('n_ffnn', C.c_int32), ('ffnn_units', C.c_int32 * EMB_MAX_FFNN), ('ffnn_dropout', C.c_float * EMB_MAX_FFNN),
                ('n_cnn', C.c_int32), ('cnn_channels', C.c_int32 * EMB_MAX_CNN), ('cnn_kernels', C.c_int32 * EMB_MAX_CNN),
                ('cnn_dropout', C.c_float * EMB_MAX_CNN),
                ('embracement_size', C.c_int32),
                ('n_post', C.c_int32), ('post_units', C.c_int32 * EMB_MAX_POST), ('post_dropout', C.c_float * EMB_MAX_POST),
                ('p_ffnn', C.c_double), ('embracenet_dropout', C.c_int32), ('reserved', C.c_int32)]


class EmbDraws(C.Structure):
    _fields_ = [('ffnn_drop', C.c_void_p * EMB_MAX_FFNN), ('cnn_drop', C.c_void_p * EMB_MAX_CNN),
                ('post_drop', C.c_void_p * EMB_MAX_POST), ('embrace_u', C.c_void_p), ('modal_rows', C.c_void_p),
                ('modal_u0', C.c_float), ('has_modal_u0', C.c_int32)]


class EmbOptConfig(C.Structure):
    _fields_ = [('kind', C.c_int32), ('lr', C.c_float), ('weight_decay', C.c_float), ('beta1', C.c_float),
                ('beta2', C.c_float), ('eps', C.c_float), ('alpha', C.c_float), ('momentum_decay', C.c_float)]


class EmbParamInfo(C.Structure):
    _fields_ = [('name', C.c_char * 64), ('offset', C.c_int64), ('numel', C.c_int64), ('ndim', C.c_int32),
                ('shape', C.c_int32 * 3), ('is_buffer', C.c_int32)]


class EmbStepMetrics(C.Structure):
    _fields_ = [('loss', C.c_float), ('tp', C.c_int32), ('fp', C.c_int32), ('fn', C.c_int32), ('tn', C.c_int32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
PHASE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p)

# every symbol include/embrace_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    'emb_last_error': (C.c_char_p, []),
    'emb_abi_version': (C.c_int, []),
    'emb_device_count': (C.c_int, []),
    'emb_set_option': (C.c_int, [C.c_char_p, C.c_int32]),
    'emb_get_option': (C.c_int, [C.c_char_p, C.POINTER(C.c_int32)]),
    'emb_create': (C.c_int, [C.POINTER(EmbArchSpec), C.c_int32, C.c_int32, C.POINTER(_P)]),
    'emb_destroy': (None, [_P]),
    'emb_param_count': (C.c_int64, [_P]),
    'emb_buffer_count': (C.c_int64, [_P]),
    'emb_workspace_bytes': (C.c_int64, [_P]),
    'emb_num_tensors': (C.c_int32, [_P]),
    'emb_param_info': (C.c_int, [_P, C.c_int32, C.POINTER(EmbParamInfo)]),
    'emb_output_size': (C.c_int32, [_P, C.c_int32]),
    'emb_bind': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64]),
    'emb_set_seed': (C.c_int, [_P, C.c_uint64]),
    'emb_set_shard': (C.c_int, [_P, C.c_int64, C.c_int64]),
    'emb_set_global_positives': (C.c_int, [_P, C.c_int64]),
    'emb_forward_train': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(EmbDraws), _P, _P]),
    'emb_forward_infer': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(EmbDraws), _P, _P, _P]),
    'emb_loss_ce_weighted': (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    'emb_backward': (C.c_int, [_P, _P, _P]),
    'emb_opt_step': (C.c_int, [_P, C.POINTER(EmbOptConfig), _P]),
    'emb_opt_state_get': (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    'emb_opt_state_set': (C.c_int, [_P, C.c_int64, C.c_double]),
    'emb_train_step': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(EmbDraws), C.POINTER(EmbOptConfig), _P]),
    'emb_train_step_host': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(EmbOptConfig), C.POINTER(EmbStepMetrics), _P]),
    'emb_train_step_host_pipelined': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(EmbOptConfig), C.POINTER(EmbStepMetrics), C.POINTER(C.c_int32), _P]),
    'emb_train_step_host_flush': (C.c_int, [_P, C.POINTER(EmbStepMetrics), C.POINTER(C.c_int32), _P]),
    'emb_predict_host': (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P]),
    'emb_predict_host_pipelined': (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, C.POINTER(C.c_int32), _P]),
    'emb_predict_host_flush': (C.c_int, [_P, _P]),
    'emb_metrics_reset': (C.c_int, [_P, _P]),
    'emb_metrics_read': (C.c_int, [_P, C.POINTER(EmbStepMetrics), C.c_int32, _P]),
    'emb_last_selection': (C.c_int, [_P, _P, C.c_int32, _P]),
    'emb_launch_count': (C.c_int64, [_P]),
    'emb_set_tensor_core': (C.c_int, [_P, C.c_int32]),
    'emb_set_graph': (C.c_int, [_P, C.c_int32]),
    'emb_profile_gemm': (C.c_int, [_P, C.c_int32]),
    'emb_profile_read': (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    'emb_dp_comm_bytes': (C.c_int64, []),
    'emb_dp_attach': (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    'emb_dp_detach': (C.c_int, [_P]),
    'emb_dp_allreduce_grads': (C.c_int, [_P, _P]),
    'emb_ipc_export': (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    'emb_ipc_open': (C.c_int, [_P, C.c_int64, C.POINTER(_P)]),
    'emb_set_allreduce': (C.c_int, [_P, ALLREDUCE_FN, _P]),
    'emb_set_phase_hook': (C.c_int, [_P, PHASE_FN, _P]),
    'emb_k_onehot_conv_fwd': (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    'emb_k_onehot_conv_bwd': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    'emb_k_onehot_conv_fwd_tc': (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    'emb_k_onehot_conv_wgrad_tc': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    'emb_k_umma_shift_probe': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    'emb_k_gemm': (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P] + [C.c_int32] * 8 + [_P]),
    'emb_k_gemm_time': (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P] + [C.c_int32] * 9 + [C.POINTER(C.c_float), _P]),
}

_lib = None


class EmbError(RuntimeError):
    pass


def sources():
    return [os.path.join(CSRC, 'engine.cu')]


def build(force=False, verbose=False):
    """Compile libembrace_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = sources()
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, 'embrace_b200.h')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ['nvcc'] + NVCC_FLAGS + ['-o', LIB_PATH] + srcs + ['-lcuda']
    if verbose:
        print(' '.join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise EmbError('nvcc failed:\n' + r.stdout + r.stderr)
    return LIB_PATH


def lib():
    """The loaded library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmbError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(the engine has no CPU or PyTorch fallback)')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def set_option(name, value):
    """Process-wide tuning switch of the library (include/embrace_b200.h: emb_set_option)."""
    check(lib().emb_set_option(name.encode(), int(value)))


def get_option(name):
    v = C.c_int32()
    check(lib().emb_get_option(name.encode(), C.byref(v)))
    return v.value


def check(rc):
    if rc < 0:
        raise EmbError(f'libembrace_sm100 error {rc}: {lib().emb_last_error().decode()}')
    return rc
