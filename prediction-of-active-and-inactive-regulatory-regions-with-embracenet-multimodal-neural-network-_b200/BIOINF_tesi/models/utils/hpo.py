"""A small Optuna-compatible substrate for the hyper-parameter sweeps.

optuna is not installed in the target image, and the sweep is trial-parallel across GPUs, so the pieces the reference
uses (training_models_multimodal.py:264-276, :420-437) are restated here with the same names and defaults:
  Trial            suggest_int / suggest_categorical / suggest_float / suggest_loguniform / report / should_prune / number / params
  RandomSampler, TPESampler (independent Parzen estimators; n_startup_trials = 10 random trials first, like Optuna's)
  MedianPruner (n_startup_trials = 5, n_warmup_steps = 0), PatientPruner(wrapped, patience)
  create_study(study_name, direction, pruner, storage, load_if_exists, sampler) -> Study.optimize(objective, n_trials)
Storage is an append-only JSON-lines file guarded by flock, shared by the worker processes of a sweep (one per GPU);
`load_if_exists` resumes a study exactly as the reference's sqlite storage does.
"""
import fcntl
import json
import math
import os
import random
import time


class TrialPruned(Exception):
    pass


class TrialState:
    RUNNING, COMPLETE, PRUNED, FAIL = 'RUNNING', 'COMPLETE', 'PRUNED', 'FAIL'


# ---- distributions ---------------------------------------------------------------------------------
def _dist_int(lo, hi):
    return {'type': 'int', 'low': int(lo), 'high': int(hi)}


def _dist_cat(choices):
    return {'type': 'cat', 'choices': list(choices)}


def _dist_float(lo, hi, log):
    return {'type': 'float', 'low': float(lo), 'high': float(hi), 'log': bool(log)}


class FrozenTrial:
    def __init__(self, number, state=TrialState.RUNNING, params=None, distributions=None, value=None, intermediate=None,
                 datetime_start=None, datetime_complete=None):
        self.number, self.state, self.value = number, state, value
        self.params = dict(params or {})
        self.distributions = dict(distributions or {})
        self.intermediate_values = {int(k): v for k, v in (intermediate or {}).items()}
        self.datetime_start, self.datetime_complete = datetime_start, datetime_complete

    def to_json(self):
        return dict(number=self.number, state=self.state, value=self.value, params=self.params, distributions=self.distributions,
                    intermediate=self.intermediate_values, datetime_start=self.datetime_start, datetime_complete=self.datetime_complete)


# ---- samplers ----------------------------------------------------------------------------------------
class RandomSampler:
    def __init__(self, seed=None):
        self.rng = random.Random(seed)

    def sample(self, study, trial, name, dist):
        return _sample_uniform(self.rng, dist)


def _sample_uniform(rng, dist):
    if dist['type'] == 'int':
        return rng.randint(dist['low'], dist['high'])
    if dist['type'] == 'cat':
        return dist['choices'][rng.randrange(len(dist['choices']))]
    if dist['log']:
        return min(max(math.exp(rng.uniform(math.log(dist['low']), math.log(dist['high']))), dist['low']), dist['high'])
    return rng.uniform(dist['low'], dist['high'])


class TPESampler:
    """Tree-structured Parzen estimator, univariate per parameter: after `n_startup_trials` random trials the finished
    trials are split at the gamma-quantile of the objective into 'good' and 'bad'; `n_ei_candidates` values are drawn
    from the good density l(x) and the one maximising l(x) / g(x) is kept."""

    def __init__(self, seed=None, n_startup_trials=10, n_ei_candidates=24, gamma=0.1):
        self.rng = random.Random(seed)
        self.n_startup_trials, self.n_ei_candidates, self.gamma = n_startup_trials, n_ei_candidates, gamma

    def sample(self, study, trial, name, dist):
        done = [t for t in study.trials if t.state == TrialState.COMPLETE and name in t.params and t.value is not None]
        if len(done) < self.n_startup_trials:
            return _sample_uniform(self.rng, dist)
        sign = -1.0 if study.direction == 'maximize' else 1.0
        done.sort(key=lambda t: sign * t.value)
        n_good = max(1, min(int(math.ceil(self.gamma * len(done))), 25))
        good, bad = done[:n_good], done[n_good:] or done[:n_good]
        if dist['type'] == 'cat':
            ch = dist['choices']

            def probs(ts):
                w = [1.0] * len(ch)                      # Laplace prior
                for t in ts:
                    if t.params[name] in ch:
                        w[ch.index(t.params[name])] += 1.0
                s = sum(w)
                return [x / s for x in w]
            pl, pg = probs(good), probs(bad)
            cands = self.rng.choices(range(len(ch)), weights=pl, k=self.n_ei_candidates)
            return ch[max(cands, key=lambda i: pl[i] / pg[i])]
        log = dist['type'] == 'float' and dist['log']
        f = (lambda v: math.log(v)) if log else (lambda v: float(v))
        lo, hi = f(dist['low']), f(dist['high'])
        width = max(hi - lo, 1e-12)

        def parzen(ts):
            mus = [f(t.params[name]) for t in ts]
            sig = max(width / max(len(mus), 1) ** 0.5 / 2.0, width * 1e-3)
            return mus + [(lo + hi) / 2.0], [sig] * len(mus) + [width]          # observations + a wide prior component

        def pdf(x, mus, sigs):
            return sum(math.exp(-0.5 * ((x - m) / s) ** 2) / s for m, s in zip(mus, sigs)) / len(mus) + 1e-300
        ml, sl = parzen(good)
        mg, sg = parzen(bad)
        best, best_score = None, -1.0
        for _ in range(self.n_ei_candidates):
            k = self.rng.randrange(len(ml))
            x = min(max(self.rng.gauss(ml[k], sl[k]), lo), hi)
            score = pdf(x, ml, sl) / pdf(x, mg, sg)
            if score > best_score:
                best, best_score = x, score
        v = min(max(math.exp(best), dist['low']), dist['high']) if log else best      # exp(log(hi)) may overshoot by an ulp
        if dist['type'] == 'int':
            v = int(min(max(round(v), dist['low']), dist['high']))
        return v


# ---- pruners -----------------------------------------------------------------------------------------
class NopPruner:
    def prune(self, study, trial):
        return False


class MedianPruner:
    """Prune when the trial's best intermediate value so far is worse than the median of the other trials' intermediate
    values at the same step (Optuna defaults: n_startup_trials = 5 finished trials first, n_warmup_steps = 0)."""

    def __init__(self, n_startup_trials=5, n_warmup_steps=0):
        self.n_startup_trials, self.n_warmup_steps = n_startup_trials, n_warmup_steps

    def prune(self, study, trial):
        if not trial.intermediate_values:
            return False
        step = max(trial.intermediate_values)
        if step < self.n_warmup_steps:
            return False
        done = [t for t in study.trials if t.state == TrialState.COMPLETE]
        if len(done) < self.n_startup_trials:
            return False
        others = sorted(t.intermediate_values[step] for t in done if step in t.intermediate_values)
        if not others:
            return False
        n = len(others)
        median = others[n // 2] if n % 2 else 0.5 * (others[n // 2 - 1] + others[n // 2])
        vals = list(trial.intermediate_values.values())
        if study.direction == 'maximize':
            return max(vals) < median
        return min(vals) > median


class PatientPruner:
    """Defers to the wrapped pruner only after the objective has failed to improve for `patience` consecutive steps."""

    def __init__(self, wrapped_pruner, patience, min_delta=0.0):
        self.wrapped, self.patience, self.min_delta = wrapped_pruner or NopPruner(), patience, min_delta

    def prune(self, study, trial):
        steps = sorted(trial.intermediate_values)
        if len(steps) <= self.patience + 1:
            return False
        vals = [trial.intermediate_values[s] for s in steps]
        before, after = vals[:-self.patience - 1], vals[-self.patience - 1:]
        if study.direction == 'maximize':
            improved = max(after) > max(before) + self.min_delta if before else True
        else:
            improved = min(after) < min(before) - self.min_delta if before else True
        if improved:
            return False
        return self.wrapped.prune(study, trial)


# ---- storage -----------------------------------------------------------------------------------------
class JsonlStorage:
    """Append-only event log: {'study', 'event': 'trial', ...FrozenTrial}.  The last record of a (study, number) wins."""

    def __init__(self, path):
        self.path = path

    def _locked(self, mode):
        f = open(self.path, mode)
        fcntl.flock(f, fcntl.LOCK_EX)
        return f

    def load(self, study_name):
        if not os.path.exists(self.path):
            return []
        trials = {}
        with self._locked('r') as f:
            for line in f:
                line = line.strip()
                if not line:
                    continue
                r = json.loads(line)
                if r.get('study') == study_name:
                    trials[r['number']] = FrozenTrial(r['number'], r['state'], r['params'], r['distributions'], r['value'],
                                                      r['intermediate'], r.get('datetime_start'), r.get('datetime_complete'))
        return [trials[k] for k in sorted(trials)]

    def next_number(self, study_name):
        """Claims the next trial number atomically (several workers may feed one study)."""
        with self._locked('a+') as f:
            f.seek(0)
            nums = [json.loads(l)['number'] for l in f if l.strip() and json.loads(l).get('study') == study_name]
            n = max(nums) + 1 if nums else 0
            f.write(json.dumps(dict(study=study_name, **FrozenTrial(n, datetime_start=time.time()).to_json())) + '\n')
        return n

    def write(self, study_name, trial):
        with self._locked('a') as f:
            f.write(json.dumps(dict(study=study_name, **trial.to_json())) + '\n')


class MemoryStorage:
    def __init__(self):
        self.data = {}

    def load(self, study_name):
        return [self.data[study_name][k] for k in sorted(self.data.get(study_name, {}))]

    def next_number(self, study_name):
        d = self.data.setdefault(study_name, {})
        n = max(d) + 1 if d else 0
        d[n] = FrozenTrial(n, datetime_start=time.time())
        return n

    def write(self, study_name, trial):
        self.data.setdefault(study_name, {})[trial.number] = trial


# ---- trial / study -----------------------------------------------------------------------------------
class Trial:
    """What the model constructors and objectives see (duck-typed like optuna.trial.Trial)."""

    def __init__(self, study, number):
        self.study, self.number = study, number
        self._t = FrozenTrial(number, datetime_start=time.time())

    @property
    def params(self):
        return self._t.params

    def _suggest(self, name, dist):
        if name in self._t.params:
            return self._t.params[name]
        fixed = self.study.fixed_params
        v = fixed[name] if fixed and name in fixed else self.study.sampler.sample(self.study, self._t, name, dist)
        self._t.params[name] = v
        self._t.distributions[name] = dist
        return v

    def suggest_int(self, name, low, high):
        return self._suggest(name, _dist_int(low, high))

    def suggest_categorical(self, name, choices):
        return self._suggest(name, _dist_cat(choices))

    def suggest_float(self, name, low, high, log=False):
        return self._suggest(name, _dist_float(low, high, log))

    def suggest_uniform(self, name, low, high):
        return self._suggest(name, _dist_float(low, high, False))

    def suggest_loguniform(self, name, low, high):
        return self._suggest(name, _dist_float(low, high, True))

    def report(self, value, step):
        self._t.intermediate_values[int(step)] = float(value)
        self.study.storage.write(self.study.study_name, self._t)

    def should_prune(self):
        return self.study.pruner.prune(self.study, self._t)


class Study:
    def __init__(self, study_name, direction, sampler, pruner, storage):
        self.study_name, self.direction = study_name, direction
        self.sampler, self.pruner, self.storage = sampler or TPESampler(), pruner or MedianPruner(), storage
        self.fixed_params = None

    @property
    def trials(self):
        return self.storage.load(self.study_name)

    @property
    def best_trial(self):
        done = [t for t in self.trials if t.state == TrialState.COMPLETE and t.value is not None]
        if not done:
            raise ValueError('no completed trial')
        return (max if self.direction == 'maximize' else min)(done, key=lambda t: t.value)

    @property
    def best_params(self):
        return self.best_trial.params

    @property
    def best_value(self):
        return self.best_trial.value

    def enqueue_trial(self, params):
        self.fixed_params = dict(params)

    def optimize(self, objective, n_trials):
        for _ in range(n_trials):
            trial = Trial(self, self.storage.next_number(self.study_name))
            try:
                value = objective(trial)
                trial._t.state, trial._t.value = TrialState.COMPLETE, float(value)
            except TrialPruned:
                trial._t.state = TrialState.PRUNED
                iv = trial._t.intermediate_values
                trial._t.value = iv[max(iv)] if iv else None
            except Exception:
                trial._t.state = TrialState.FAIL
                trial._t.datetime_complete = time.time()
                self.storage.write(self.study_name, trial._t)
                raise
            trial._t.datetime_complete = time.time()
            self.storage.write(self.study_name, trial._t)
            self.fixed_params = None


def create_study(study_name=None, direction='maximize', pruner=None, storage=None, load_if_exists=False, sampler=None):
    """storage: path of a JSON-lines file (a 'sqlite:///name.db' URL maps to 'name.jsonl'), or None for in-memory."""
    if direction not in ('maximize', 'minimize'):
        raise ValueError(direction)
    if storage is None:
        st = MemoryStorage()
    else:
        path = storage
        if path.startswith('sqlite:///'):
            path = os.path.splitext(path[len('sqlite:///'):])[0] + '.jsonl'
        st = JsonlStorage(path)
    study = Study(study_name or 'study', direction, sampler, pruner, st)
    if not load_if_exists and st.load(study.study_name):
        raise ValueError(f'study {study.study_name!r} already exists in {storage} (pass load_if_exists=True)')
    return study
