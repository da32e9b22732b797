"""Shared machinery of the nn.Module mirrors: parameters are views into the engine's flat fp32 arenas."""
import ctypes as C

import torch
import torch.nn as nn

from ... import _native as N
from ...archspec import ArchSpec
from ...engine import Engine


def _resolve_device(device):
    dev = torch.device(device if device is not None else 'cuda')
    if dev.type != 'cuda':
        raise N.EmbError(f"device={device!r}: the B200 engine has no CPU path (the reference's CPU route is the oracle's job)")
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return dev


class _EngineFn(torch.autograd.Function):
    """logits = engine.forward_train(...); backward hands dlogits to emb_backward, which fills the gradient arena."""

    @staticmethod
    def forward(ctx, anchor, owner, x_ffnn, bases, availabilities, draws):
        ctx.owner = owner
        logits = owner._engine_for(x_ffnn, bases).forward(x_ffnn, bases, training=True, draws=draws, availabilities=availabilities)
        ctx.token = owner._forward_token
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        owner = ctx.owner
        if ctx.token != owner._forward_token:
            raise RuntimeError('backward() through a stale forward: the engine keeps the activations of the most recent '
                               'training forward only')
        owner._engine.backward(dlogits.contiguous())
        owner._publish_grads()
        return (None,) * 6


class EngineModule(nn.Module):
    """Base of the mirrors.  Subclasses build the reference's module tree (same attribute names -> same state_dict
    keys) out of ordinary torch.nn layers used purely as PARAMETER CONTAINERS; `_adopt()` then re-points every
    parameter / BatchNorm buffer at its slot in the engine's arenas.  No torch.nn forward is ever executed."""

    precision_default = 'bf16'

    def _adopt(self, spec: ArchSpec, device, precision=None, seed=0x5EED):
        self.spec = spec
        self._dev = _resolve_device(device)
        self._precision = precision or self.precision_default
        self._seed = seed
        self._engine = None
        self._forward_token = 0
        self._anchor = None
        lib = N.lib()
        if not torch.cuda.is_available() or lib.emb_device_count() < 1:
            raise N.EmbError('no B200 (sm_100) device visible: the EmbraceNet engine has no CPU fallback')
        # plan the arenas (no device work) and create them on the Python side so that they outlive workspace resizes
        h = C.c_void_p()
        N.check(lib.emb_create(C.byref(spec.to_c()), 1, N.PREC[self._precision], C.byref(h)))
        n_params, n_buffers = lib.emb_param_count(h), lib.emb_buffer_count(h)
        table, info = [], N.EmbParamInfo()
        for i in range(lib.emb_num_tensors(h)):
            N.check(lib.emb_param_info(h, i, C.byref(info)))
            table.append((info.name.decode(), info.offset, info.numel, tuple(info.shape[:info.ndim]), bool(info.is_buffer)))
        lib.emb_destroy(h)
        f32 = dict(dtype=torch.float32, device=self._dev)
        self._arenas = dict(params=torch.zeros(max(n_params, 4), **f32), grads=torch.zeros(max(n_params, 4), **f32),
                            buffers=torch.zeros(max(n_buffers, 4), **f32), opt_m=torch.zeros(max(n_params, 4), **f32),
                            opt_v=torch.zeros(max(n_params, 4), **f32))
        self._table = table
        own = dict(self.named_parameters())
        bufs = dict(self.named_buffers())
        with torch.no_grad():
            for name, off, numel, shape, is_buf in table:
                arena = self._arenas['buffers' if is_buf else 'params']
                view = arena[off:off + numel].view(shape)
                src = bufs[name] if is_buf else own[name]
                assert tuple(src.shape) == shape, (name, tuple(src.shape), shape)
                view.copy_(src.detach().to(self._dev, torch.float32))
                if is_buf:
                    mod, leaf = self._owner_of(name)
                    mod._buffers[leaf] = view
                else:
                    own[name].data = view
        for b in self.buffers():       # num_batches_tracked counters follow the arenas to the device
            if b.device != self._dev:
                b.data = b.data.to(self._dev)
        self._anchor = torch.zeros((), device=self._dev, requires_grad=True)
        return self

    def _owner_of(self, dotted):
        mod = self
        parts = dotted.split('.')
        for p in parts[:-1]:
            mod = getattr(mod, p)
        return mod, parts[-1]

    # ---- nn.Module API the reference's L3/L4 code uses ---------------------------------------------------------
    def double(self):
        """The reference trains `model.double()` on CPU; the engine keeps fp32 master weights and computes in
        bf16 (or fp32): the call is accepted and ignored (precision contract)."""
        return self

    def float(self):
        return self

    def to(self, *args, **kwargs):
        dev = kwargs.get('device', args[0] if args and not isinstance(args[0], torch.dtype) else None)
        if dev is not None and torch.device(dev).type != 'cuda':
            raise N.EmbError('the B200 engine cannot be moved off the GPU')
        return self

    def cuda(self, device=None):
        return self

    def _apply(self, fn, recurse=True):
        # parameters are arena views: dtype / device conversions must not re-allocate them
        return self

    def _engine_for(self, x_ffnn, bases):
        """The engine whose workspace holds a batch of this size.  The first engine is sized with 12.5 % + 8 rows of slack
        (BalancePos_BatchSampler batches differ by a row or two and arrive shuffled).  A batch that still does not fit
        re-creates the engine over the same parameter / optimizer arenas: the metric records of the running epoch are
        carried over on the host (metrics_read() prepends them), the optimizer step count is kept, and the new engine's
        generator is keyed by a seed derived from the resize count, so its dropout / selection draws do not replay the
        old engine's."""
        B = (x_ffnn if x_ffnn is not None else bases).shape[0]
        if self._engine is None or B > self._engine.max_batch:
            old = self._engine
            cap = max(B + B // 8 + 8, 256)
            state, seed = None, self._seed
            if old is not None:
                state = old.opt_state()
                self._carry_records = getattr(self, '_carry_records', []) + old.metrics_read()
                self._resizes = getattr(self, '_resizes', 0) + 1
                seed = (self._seed + 0x9E3779B97F4A7C15 * self._resizes) & 0xFFFFFFFFFFFFFFFF
            self._engine = Engine(self.spec, cap, precision=self._precision, device=self._dev, seed=seed, arenas=self._arenas)
            if state is not None:
                self._engine.set_opt_state(*state)
            del old
        return self._engine

    def metrics_reset(self):
        self._carry_records = []
        if self._engine is not None:
            self._engine.metrics_reset()

    def metrics_read(self):
        """Per-batch records since the last reset, including those an engine resize carried over."""
        rec = list(getattr(self, '_carry_records', []))
        return rec + (self._engine.metrics_read() if self._engine is not None else [])

    def _publish_grads(self):
        own = dict(self.named_parameters())
        for name, off, numel, shape, is_buf in self._table:
            if not is_buf:
                own[name].grad = self._arenas['grads'][off:off + numel].view(shape)

    def _bump_batches_tracked(self):
        # the modules that own a counter are looked up once (walking named_buffers() on every step was ~90 us of host time per step);
        # the tensor itself is fetched from the module each time, so .to() / .double() / load_state_dict stay safe
        owners = self.__dict__.get('_nbt_owners')
        if owners is None:
            owners = [m for m in self.modules() if 'num_batches_tracked' in m._buffers]
            self.__dict__['_nbt_owners'] = owners
        for m in owners:
            m._buffers['num_batches_tracked'] += 1

    @staticmethod
    def _to_bases(x_cnn, dev):
        """[B,4,256] one-hot (any float dtype) or [B,256] integer codes (a,c,g,t = 0..3) -> uint8 [B,256] on the device."""
        x = torch.as_tensor(x_cnn)
        if x.dim() == 2 and x.shape[1] == 256:
            return x.to(dev, torch.uint8)
        if x.dim() == 3 and x.shape[1] == 4 and x.shape[2] == 256:
            return x.to(dev).argmax(dim=1).to(torch.uint8)
        raise ValueError(f'sequence input must be one-hot [B,4,256] or base codes [B,256], got {tuple(x.shape)}')

    def _run(self, x_ffnn, x_cnn, availabilities, draws, modality_dropout):
        dev = self._dev
        xf = torch.as_tensor(x_ffnn).to(dev, torch.float32).contiguous() if x_ffnn is not None else None
        if xf is not None and xf.dim() == 1:
            xf = xf.unsqueeze(0)
        bases = self._to_bases(x_cnn, dev).contiguous() if x_cnn is not None else None
        if self.spec.kind == 'embracenet' and not modality_dropout:
            draws = dict(draws or {})
            draws.setdefault('modal_u0', 0.0)          # coin < 0.5: EmbraceNetMultimodal.py:179-180 takes no modality away
        if self.training:
            self._forward_token += 1
            if torch.is_grad_enabled():
                out = _EngineFn.apply(self._anchor, self, xf, bases, availabilities, draws)
            else:
                out = self._engine_for(xf, bases).forward(xf, bases, training=True, draws=draws, availabilities=availabilities)
            self._bump_batches_tracked()
            return out
        return self._engine_for(xf, bases).forward(xf, bases, training=False, draws=draws, availabilities=availabilities)

    def _batch_inputs(self, x_ffnn, x_cnn, target):
        dev = self._dev
        xf = torch.as_tensor(x_ffnn).to(dev, torch.float32).contiguous() if x_ffnn is not None else None
        bases = self._to_bases(x_cnn, dev).contiguous() if x_cnn is not None else None
        y = torch.as_tensor(target).reshape(-1).to(dev, torch.int32).contiguous()
        return xf, bases, y

    def train_batch(self, x_ffnn, x_cnn, target, opt_cfg, draws=None, reset_metrics=False):
        """One iteration of the reference's loop body (training_models_multimodal.py:132-162) as a single fused engine
        call: forward(is_training=True), class-weighted CE, backward, optimizer step; loss and confusion counts are
        appended to the device-side metric records."""
        xf, bases, y = self._batch_inputs(x_ffnn, x_cnn, target)
        if reset_metrics:
            self.metrics_reset()
        eng = self._engine_for(xf, bases)
        self._forward_token += 1
        eng.train_step(xf, bases, y, opt_cfg, draws=draws)
        self._bump_batches_tracked()

    @torch.no_grad()
    def eval_batch(self, x_ffnn, x_cnn, target, draws=None, reset_metrics=False):
        """model.eval() forward + the same loss/metric record, no gradient (training_models_multimodal.py:167-192)."""
        xf, bases, y = self._batch_inputs(x_ffnn, x_cnn, target)
        if reset_metrics:
            self.metrics_reset()
        eng = self._engine_for(xf, bases)
        logits = eng.forward(xf, bases, training=False, draws=draws)
        eng.loss(logits, y, want_grad=False)
        return logits

    @torch.no_grad()
    def predict_scores(self, x_FFNN=None, x_CNN=None, availabilities=None, batch_size=65536, draws=None, column='prob'):
        """Batched form of the predict loop `[model_(x_i)[1] for i in range(n)]` (visual.py:284-293): one eval-mode forward
        per `batch_size` rows.  column='prob' returns softmax(logits)[:, 1] (the `_NoTrain` twins that apply their
        softmax), column='logit' returns logits[:, 1] (ConcatNetMultimodal_NoTrain, whose softmax is lost to a typo).
        `draws['embrace_u']`, when given, holds one row of uniforms per sample and is sliced with the batch."""
        was_training = self.training
        self.eval()
        out = []
        n = len(x_FFNN if x_FFNN is not None else x_CNN)
        for lo in range(0, n, batch_size):
            hi = min(n, lo + batch_size)
            xf = torch.as_tensor(x_FFNN[lo:hi]).to(self._dev, torch.float32).contiguous() if x_FFNN is not None else None
            bases = self._to_bases(x_CNN[lo:hi], self._dev).contiguous() if x_CNN is not None else None
            av = None if availabilities is None else torch.as_tensor(availabilities[lo:hi])
            d = None if draws is None else {k: (v[lo:hi] if k == 'embrace_u' else v) for k, v in draws.items()}
            if self.spec.kind == 'embracenet':
                d = dict(d or {})
                d.setdefault('modal_u0', 0.0)
            logits, probs = self._engine_for(xf, bases).forward(xf, bases, training=False, draws=d, availabilities=av, want_probs=True)
            out.append(probs if column == 'prob' else logits[:, 1].clone())
        self.train(was_training)
        return torch.cat(out)

    def state_dict(self, *args, **kwargs):
        """Reference key names; tensors are detached COPIES (the live parameters are views of one flat arena)."""
        sd = super().state_dict(*args, **kwargs)
        return type(sd)((k, v.detach().clone()) for k, v in sd.items())

    def __getstate__(self):
        raise TypeError('pickling a whole engine-backed module is not supported: save model.state_dict() (the reference\'s '
                        'torch.save(model) of Optuna trials, training_models_multimodal.py:413, is outside the hot path)')

    @property
    def engine(self):
        return self._engine
