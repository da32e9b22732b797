"""Host-side owner of one libembrace_sm100 engine.

PyTorch is plumbing here: it owns the device memory (flat fp32 parameter / gradient / optimizer arenas
and the workspace) and the stream; every computation is a call through the C ABI with raw device
pointers.  There is deliberately no fallback: no GPU or no built library -> EmbError.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from . import _native as N
from .archspec import ArchSpec


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    def __init__(self, spec: ArchSpec, max_batch: int, precision: str = 'fp32', device='cuda', seed: int = 0x5EED,
                 tensor_core: bool = None, arenas=None):
        """arenas: optional dict(params=, grads=, buffers=, opt_m=, opt_v=) of fp32 CUDA tensors to bind instead of
        allocating new ones (the nn.Module mirror owns them so that its nn.Parameters survive a workspace resize)."""
        self.lib = N.lib()
        if not torch.cuda.is_available() or self.lib.emb_device_count() < 1:
            raise N.EmbError('no B200 (sm_100) device visible: the EmbraceNet engine has no CPU fallback')
        self.spec = spec
        self.max_batch = int(max_batch)
        self.precision = precision
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise N.EmbError('the engine runs on CUDA devices only')
        self._h = C.c_void_p()
        cspec = spec.to_c()
        N.check(self.lib.emb_create(C.byref(cspec), self.max_batch, N.PREC[precision], C.byref(self._h)))
        self.n_params = self.lib.emb_param_count(self._h)
        self.n_buffers = self.lib.emb_buffer_count(self._h)
        self.ws_bytes = self.lib.emb_workspace_bytes(self._h)
        with torch.cuda.device(self.device):
            f32 = dict(dtype=torch.float32, device=self.device)
            arenas = arenas or {}
            self.params = arenas.get('params', None)
            if self.params is None:
                self.params = torch.zeros(max(self.n_params, 4), **f32)
            self.grads = arenas['grads'] if 'grads' in arenas else torch.zeros(max(self.n_params, 4), **f32)
            self.buffers = arenas['buffers'] if 'buffers' in arenas else torch.zeros(max(self.n_buffers, 4), **f32)
            self.opt_m = arenas['opt_m'] if 'opt_m' in arenas else torch.zeros(max(self.n_params, 4), **f32)
            self.opt_v = arenas['opt_v'] if 'opt_v' in arenas else torch.zeros(max(self.n_params, 4), **f32)
            for t in (self.params, self.grads, self.buffers, self.opt_m, self.opt_v):
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
            self.workspace = torch.zeros(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 256
            self._ws_view = self.workspace[off:off + self.ws_bytes]
            N.check(self.lib.emb_bind(self._h, _ptr(self.params), _ptr(self.grads), _ptr(self.buffers), _ptr(self.opt_m),
                                      _ptr(self.opt_v), _ptr(self._ws_view), self.ws_bytes))
        N.check(self.lib.emb_set_seed(self._h, seed))
        self.table = []
        info = N.EmbParamInfo()
        for i in range(self.lib.emb_num_tensors(self._h)):
            N.check(self.lib.emb_param_info(self._h, i, C.byref(info)))
            self.table.append(dict(name=info.name.decode(), offset=info.offset, numel=info.numel,
                                   shape=tuple(info.shape[:info.ndim]), is_buffer=bool(info.is_buffer)))
        if tensor_core is None:
            tensor_core = precision == 'bf16'
        self.set_tensor_core(tensor_core)
        self._keep = None
        self._ext_streams = {}

    def __del__(self):
        try:
            if getattr(self, '_h', None) and self._h.value:
                self.lib.emb_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------------------
    def view(self, entry, which='params'):
        """Tensor view (reference state_dict shape) of one table entry in the params / grads / buffers arena."""
        arena = self.buffers if entry['is_buffer'] else getattr(self, which)
        return arena[entry['offset']:entry['offset'] + entry['numel']].view(entry['shape'])

    def named_views(self, which='params'):
        return {t['name']: self.view(t, which) for t in self.table if which == 'params' or not t['is_buffer']}

    def load_numpy(self, P):
        """P: {state_dict key: numpy array}; num_batches_tracked entries are ignored (host-side counter)."""
        for t in self.table:
            self.view(t).copy_(torch.from_numpy(np.asarray(P[t['name']], dtype=np.float32)).to(self.device))

    def init_random(self, seed=0):
        """PyTorch-default-style initialisation in the arenas: nn.Linear / nn.Conv1d weights and biases ~ U(+-1/sqrt(fan_in)),
        BatchNorm weight 1, bias 0, running_mean 0, running_var 1 (the state a freshly constructed reference model has)."""
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        shapes = {t['name']: t['shape'] for t in self.table}
        for t in self.table:
            name, shape = t['name'], t['shape']
            module = name.rsplit('.', 1)[0]
            if name.endswith('running_mean'):
                v = torch.zeros(shape)
            elif name.endswith('running_var'):
                v = torch.ones(shape)
            elif module + '.running_mean' in shapes:
                v = torch.ones(shape) if name.endswith('weight') else torch.zeros(shape)
            else:
                w = shapes[module + '.weight']
                bound = 1.0 / float(np.sqrt(max(int(np.prod(w[1:])), 1)))
                v = (torch.rand(shape, generator=g) * 2 - 1) * bound
            self.view(t).copy_(v.to(self.device))

    def grads_numpy(self):
        return {t['name']: self.view(t, 'grads').detach().cpu().numpy().astype(np.float64) for t in self.table if not t['is_buffer']}

    def params_numpy(self):
        return {t['name']: self.view(t).detach().cpu().numpy().astype(np.float64) for t in self.table}

    def set_seed(self, seed):
        """Re-key the counter-based generator and rewind its step counter."""
        N.check(self.lib.emb_set_seed(self._h, int(seed)))

    def set_tensor_core(self, on):
        N.check(self.lib.emb_set_tensor_core(self._h, 1 if on else 0))
        self.tensor_core = bool(on)

    def set_graph(self, on=True, collectives=False):
        """Replay the whole train step as one CUDA graph (captured the second time a batch size is seen).
        collectives=True (data parallel, opt-in): the all-reduce / phase callbacks are invoked DURING capture with torch's
        current stream switched to the capture stream, so the NCCL collectives they enqueue become nodes of the graph."""
        N.check(self.lib.emb_set_graph(self._h, (2 if collectives else 1) if on else 0))
        self.graph = bool(on)

    def set_shard(self, row_offset, global_batch):
        N.check(self.lib.emb_set_shard(self._h, int(row_offset), int(global_batch)))

    def set_global_positives(self, n_pos):
        N.check(self.lib.emb_set_global_positives(self._h, int(n_pos)))

    def _on_stream(self, stream):
        """torch's current stream := the stream the engine is enqueuing on (its own capture stream in graph mode)."""
        stream = int(stream or 0)
        if stream == torch.cuda.current_stream(self.device).cuda_stream:
            return contextlib.nullcontext()
        ext = self._ext_streams.get(stream)
        if ext is None:
            ext = self._ext_streams[stream] = torch.cuda.ExternalStream(stream, device=self.device)
        return torch.cuda.stream(ext)

    def set_allreduce(self, fn):
        """fn(tensor_float64) -> None performs an in-place SUM all-reduce of a device tensor (SyncBN statistics)."""
        base = self._ws_view.data_ptr()
        views = {}                     # (offset, count) -> float64 view: the same few buffers come back every step

        def cb(user, ptr, count, stream):
            try:
                key = (ptr - base, count)
                t = views.get(key)
                if t is None:
                    t = views[key] = self._ws_view[key[0]:key[0] + count * 8].view(torch.float64)
                with self._on_stream(stream):
                    fn(t)
                return 0
            except Exception as ex:           # never unwind through C
                print('allreduce callback failed:', ex)
                return 1
        self._allreduce_cb = N.ALLREDUCE_FN(cb) if fn is not None else N.ALLREDUCE_FN()
        N.check(self.lib.emb_set_allreduce(self._h, self._allreduce_cb, None))

    def set_phase_hook(self, fn):
        """fn(phase) -> None is called from inside backward once the non-CNN gradients are final (phase 1)."""
        def cb(user, phase, stream):
            try:
                with self._on_stream(stream):
                    fn(int(phase))
                return 0
            except Exception as ex:           # never unwind through C
                print('phase hook failed:', ex)
                return 1
        self._phase_cb = N.PHASE_FN(cb) if fn is not None else N.PHASE_FN()
        N.check(self.lib.emb_set_phase_hook(self._h, self._phase_cb, None))

    def arena_range(self, prefix):
        """[lo, hi) of the parameter-arena elements whose state_dict key starts with `prefix` (tensors are laid out in
        state_dict order, so a module's parameters are contiguous)."""
        ts = [t for t in self.table if not t['is_buffer'] and t['name'].startswith(prefix)]
        if not ts:
            return 0, 0
        return min(t['offset'] for t in ts), max(t['offset'] + t['numel'] for t in ts)

    # ---- draws -----------------------------------------------------------------------------------
    def _draws(self, draws):
        """dict of replayed uniforms (numpy or torch, reference layouts) -> EmbDraws (device pointers)."""
        if draws is None:
            self._keep = None
            return None
        d = N.EmbDraws()
        keep = []

        def dev(a, dtype):
            t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(device=self.device, dtype=dtype).contiguous()
            keep.append(t)
            return t.data_ptr()

        for key, arr, p_list in (('ffnn_drop', d.ffnn_drop, self.spec.ffnn_dropout), ('cnn_drop', d.cnn_drop, self.spec.cnn_dropout),
                                 ('post_drop', d.post_drop, self.spec.post_dropout)):
            for i, u in enumerate(draws.get(key, []) or []):
                if u is not None and i < len(p_list) and p_list[i] > 0:
                    arr[i] = dev(u, torch.float32)
        if draws.get('embrace_u') is not None:
            d.embrace_u = dev(draws['embrace_u'], torch.float64)
        if draws.get('modal_rows') is not None:
            d.modal_rows = dev(draws['modal_rows'], torch.float32)
        if draws.get('modal_u0') is not None:
            d.modal_u0 = float(draws['modal_u0'])
            d.has_modal_u0 = 1
        self._keep = keep
        return d

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- hot path ----------------------------------------------------------------------------------
    def _inputs(self, x_ffnn, bases):
        if x_ffnn is not None:
            x_ffnn = x_ffnn.to(device=self.device, dtype=torch.float32).contiguous()
        if bases is not None:
            bases = bases.to(device=self.device, dtype=torch.uint8).contiguous()
        return x_ffnn, bases

    def forward(self, x_ffnn, bases, training, draws=None, availabilities=None, want_probs=False):
        x_ffnn, bases = self._inputs(x_ffnn, bases)
        B = (x_ffnn if x_ffnn is not None else bases).shape[0]
        logits = torch.empty(B, 2, dtype=torch.float32, device=self.device)
        av = availabilities.to(device=self.device, dtype=torch.float32).contiguous() if availabilities is not None else None
        d = self._draws(draws)
        dref = C.byref(d) if d is not None else None
        if training:
            N.check(self.lib.emb_forward_train(self._h, _ptr(x_ffnn), _ptr(bases), _ptr(av), B, dref, _ptr(logits), self.stream))
            return logits
        probs = torch.empty(B, dtype=torch.float32, device=self.device) if want_probs else None
        N.check(self.lib.emb_forward_infer(self._h, _ptr(x_ffnn), _ptr(bases), _ptr(av), B, dref, _ptr(logits), _ptr(probs), self.stream))
        return (logits, probs) if want_probs else logits

    def infer(self, x_ffnn, bases, availabilities, probs_out):
        """Eval-mode forward of device-resident rows straight into `probs_out` ([B] fp32 device tensor): no allocation, no copy."""
        B = (x_ffnn if x_ffnn is not None else bases).shape[0]
        N.check(self.lib.emb_forward_infer(self._h, _ptr(x_ffnn), _ptr(bases), _ptr(availabilities), B, None, None, _ptr(probs_out), self.stream))
        return probs_out

    def loss(self, logits, labels, want_grad=True):
        B = logits.shape[0]
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        dlogits = torch.empty(B, 2, dtype=torch.float32, device=self.device) if want_grad else None
        N.check(self.lib.emb_loss_ce_weighted(self._h, _ptr(logits.contiguous()), _ptr(labels), B, _ptr(dlogits), self.stream))
        return dlogits

    def backward(self, dlogits):
        N.check(self.lib.emb_backward(self._h, _ptr(dlogits.to(torch.float32).contiguous()), self.stream))

    @staticmethod
    def opt_config(kind='adam', lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, alpha=0.99, momentum_decay=4e-3):
        return N.EmbOptConfig(N.OPT[kind], lr, weight_decay, betas[0], betas[1], eps, alpha, momentum_decay)

    def opt_step(self, cfg):
        N.check(self.lib.emb_opt_step(self._h, C.byref(cfg), self.stream))

    def opt_state(self):
        t, mp = C.c_int64(), C.c_double()
        N.check(self.lib.emb_opt_state_get(self._h, C.byref(t), C.byref(mp)))
        return t.value, mp.value

    def set_opt_state(self, step, mu_product=1.0):
        N.check(self.lib.emb_opt_state_set(self._h, int(step), float(mu_product)))

    def train_step(self, x_ffnn, bases, labels, cfg, draws=None):
        x_ffnn, bases = self._inputs(x_ffnn, bases)
        B = (x_ffnn if x_ffnn is not None else bases).shape[0]
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        d = self._draws(draws)
        N.check(self.lib.emb_train_step(self._h, _ptr(x_ffnn), _ptr(bases), _ptr(labels), B, C.byref(d) if d is not None else None,
                                        C.byref(cfg) if cfg is not None else None, self.stream))

    def train_step_host(self, x_ffnn_host, bases_host, labels_host, cfg):
        """numpy / pinned-torch host buffers in, one EmbStepMetrics out (copies inside the call)."""
        m = N.EmbStepMetrics()
        B = (x_ffnn_host if x_ffnn_host is not None else bases_host).shape[0]

        def hp(a):
            if a is None:
                return C.c_void_p(0)
            return C.c_void_p(a.data_ptr() if torch.is_tensor(a) else a.ctypes.data)
        N.check(self.lib.emb_train_step_host(self._h, hp(x_ffnn_host), hp(bases_host), hp(labels_host), B, C.byref(cfg), C.byref(m), self.stream))
        return m

    def train_step_host_pipelined(self, x_ffnn_host, bases_host, labels_host, cfg):
        """Like train_step_host, but the copy of this batch overlaps the previous step's compute and the call returns the
        PREVIOUS step's EmbStepMetrics (None on the first call); finish with flush_host()."""
        m, have = N.EmbStepMetrics(), C.c_int32(0)
        B = (x_ffnn_host if x_ffnn_host is not None else bases_host).shape[0]

        def hp(a):
            if a is None:
                return C.c_void_p(0)
            return C.c_void_p(a.data_ptr() if torch.is_tensor(a) else a.ctypes.data)
        N.check(self.lib.emb_train_step_host_pipelined(self._h, hp(x_ffnn_host), hp(bases_host), hp(labels_host), B, C.byref(cfg),
                                                       C.byref(m), C.byref(have), self.stream))
        return m if have.value else None

    def flush_host(self):
        m, have = N.EmbStepMetrics(), C.c_int32(0)
        N.check(self.lib.emb_train_step_host_flush(self._h, C.byref(m), C.byref(have), self.stream))
        return m if have.value else None

    def predict_host(self, x_ffnn_host, bases_host, availabilities_host=None):
        B = (x_ffnn_host if x_ffnn_host is not None else bases_host).shape[0]
        out = np.empty(B, dtype=np.float32)

        def hp(a):
            if a is None:
                return C.c_void_p(0)
            return C.c_void_p(a.data_ptr() if torch.is_tensor(a) else a.ctypes.data)
        N.check(self.lib.emb_predict_host(self._h, hp(x_ffnn_host), hp(bases_host), hp(availabilities_host), B,
                                          C.c_void_p(out.ctypes.data), self.stream))
        return out

    def predict_host_pipelined(self, x_ffnn_host, bases_host, availabilities_host, out_host):
        """Scoring loop step: uploads this batch while the previous one computes and enqueues its forward; `out_host` (pinned
        float32 tensor or ndarray, [B]) holds the scores once the NEXT call -- or predict_host_flush() -- has returned.  Returns
        True when the previous call's buffer became valid.  The inputs must stay alive and unchanged until then."""
        B = (x_ffnn_host if x_ffnn_host is not None else bases_host).shape[0]

        def hp(a):
            if a is None:
                return C.c_void_p(0)
            return C.c_void_p(a.data_ptr() if torch.is_tensor(a) else a.ctypes.data)
        done = C.c_int32(0)
        N.check(self.lib.emb_predict_host_pipelined(self._h, hp(x_ffnn_host), hp(bases_host), hp(availabilities_host), B, hp(out_host),
                                                    C.byref(done), self.stream))
        return bool(done.value)

    def predict_host_flush(self):
        N.check(self.lib.emb_predict_host_flush(self._h, self.stream))

    def metrics_reset(self):
        N.check(self.lib.emb_metrics_reset(self._h, self.stream))

    def metrics_read(self, max_records=65536):
        buf = (N.EmbStepMetrics * max_records)()
        n = N.check(self.lib.emb_metrics_read(self._h, buf, max_records, self.stream))
        return [dict(loss=buf[i].loss, tp=buf[i].tp, fp=buf[i].fp, fn=buf[i].fn, tn=buf[i].tn) for i in range(n)]

    def last_selection(self, B):
        idx = torch.empty(B, self.spec.embracement_size, dtype=torch.uint8, device=self.device)
        N.check(self.lib.emb_last_selection(self._h, _ptr(idx), B, self.stream))
        return idx

    def profile_gemm(self, enable=True):
        N.check(self.lib.emb_profile_gemm(self._h, 1 if enable else 0))

    def profile_read(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        N.check(self.lib.emb_profile_read(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return dict(ms=ms.value, flops=fl.value, launches=n.value)

    @property
    def launch_count(self):
        return self.lib.emb_launch_count(self._h)
