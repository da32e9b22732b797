"""Batch wire format: packed sequences, the reference's balanced batch composition, device-resident loaders."""
import random

import numpy as np
import torch

_CODE = {'a': 0, 'c': 1, 'g': 2, 't': 3}
BASES = 'acgt'


def encode_sequences(seqs, length=256, rng=random):
    """list of DNA strings -> uint8 [N, length] codes (a,c,g,t = 0..3).

    process_sequence (data_pipe/utils.py:268-276): lower-case; every 'n' of a sequence is replaced by ONE base drawn with
    random.choice(['a','c','g','t']) for that sequence (the draw happens whether or not the sequence contains an 'n', which
    keeps the generator in step with the reference).  Row c of the reference's one-hot [4, L] is base BASES[c]."""
    out = np.empty((len(seqs), length), dtype=np.uint8)
    for i, s in enumerate(seqs):
        s = s.lower()
        if len(s) != length:
            raise ValueError(f'sequence {i} has length {len(s)}, expected {length}')
        bp = rng.choice(['a', 'c', 'g', 't'])
        try:
            out[i] = [_CODE[bp if ch == 'n' else ch] for ch in s]
        except KeyError as ex:
            raise ValueError(f'sequence {i}: unexpected character {ex}') from None
    return out


def decode_onehot(x):
    """[B, 4, L] one-hot (what the reference feeds Conv1d) -> uint8 [B, L] codes."""
    x = torch.as_tensor(x)
    return x.argmax(dim=1).to(torch.uint8)


class BalancePos_BatchSampler:
    """Index batches with the positives spread evenly (dataprepare.py:418-454), bit-for-bit:
    random.seed(random_state) on EVERY __iter__, in-place shuffles of the (persisting) index lists, np.array_split into
    n_batches + 1 chunks, negatives reversed, chunks zipped and shuffled.  __len__ is n_batches although n_batches + 1
    batches are produced -- fit_multimodal divides by len() (reference quirk 11)."""

    def __init__(self, labels, batch_size, random_state=123):
        y = np.asarray(labels).reshape(-1)
        self.pos_index = list(np.nonzero(y == 1)[0])
        self.neg_index = list(np.nonzero(y == 0)[0])
        self.random_state = random_state
        self.batch_size = batch_size
        n = len(y)
        self.n_batches = n // batch_size + (1 if n % batch_size > 0 else 0)

    def __iter__(self):
        random.seed(self.random_state)
        random.shuffle(self.pos_index)
        random.shuffle(self.neg_index)
        pos_batches = np.array_split(self.pos_index, self.n_batches + 1)
        neg_batches = np.array_split(self.neg_index, self.n_batches + 1)
        neg_batches.reverse()
        balanced = [np.concatenate((p, n)).astype(np.int64).tolist() for p, n in zip(pos_batches, neg_batches)]
        random.shuffle(balanced)
        return iter(balanced)

    def __len__(self):
        return self.n_batches


class PackedDataset:
    """features fp32 [N, F], base codes uint8 [N, 256], labels int64 [N], resident on `device`."""

    def __init__(self, features, codes, labels, device='cuda'):
        self.device = torch.device(device)
        self.x = torch.as_tensor(np.asarray(features), dtype=torch.float32).to(self.device).contiguous()
        self.codes = torch.as_tensor(np.asarray(codes), dtype=torch.uint8).to(self.device).contiguous()
        self.y = torch.as_tensor(np.asarray(labels).reshape(-1), dtype=torch.int64).to(self.device)
        self.labels_host = np.asarray(labels).reshape(-1).astype(np.int64)
        if not (len(self.x) == len(self.codes) == len(self.y)):
            raise ValueError('features, sequences and labels disagree in length')

    def __len__(self):
        return len(self.y)


class DeviceLoader:
    """One modality's view of a PackedDataset: iterating yields (x[B, F] | codes[B, 256], y[B, 1]) gathered on the device.
    Two loaders built over the SAME batch plan stay in lock-step, which is what fit_multimodal's zip() relies on."""

    def __init__(self, data: PackedDataset, modality, plan, n_len):
        self.data, self.modality, self.plan, self.n_len = data, modality, plan, n_len

    def __len__(self):
        return self.n_len

    @property
    def input_size(self):
        """Feature width without starting an iteration (utils.get_input_size reads this)."""
        return int((self.data.x if self.modality == 'FFNN' else self.data.codes).shape[1])

    def __iter__(self):
        # NOT a generator: the epoch ticket is taken when iter() is called, which is what lets _Plan see that zip() opened
        # both modality loaders back to back
        return self._batches(self.plan.open(self.modality))

    def _batches(self, ticket):
        src = self.data.x if self.modality == 'FFNN' else self.data.codes
        for ix in self.plan.batches(ticket):          # device index tensors: no host->device copy per batch
            yield src.index_select(0, ix), self.data.y.index_select(0, ix).reshape(-1, 1)


class _Plan:
    """Materialises one epoch's index batches once and serves the same list to both modality loaders.

    Two iterators share an epoch exactly when they were OPENED back to back -- iter() on one modality's loader, then iter()
    on the other's, before the first produced a batch -- which is what `zip(loader['FFNN'], loader['CNN'])` does
    (training_models_multimodal.py:136).  Any other iterator (a loader walked on its own, e.g. get_input_size over the
    reference's DataLoader, :313) gets an epoch of its own and cannot shift the two modalities against each other.  In the
    reference it does: each DataLoader owns a sampler whose index lists are shuffled in place per __iter__, so after one
    solo pass features of sample i meet the sequence of sample j while the label assert still passes (positives are paired
    with positives); that mis-pairing is deliberately NOT reproduced."""

    class _Ticket:
        __slots__ = ('modality', 'epoch', 'started')

        def __init__(self, modality):
            self.modality, self.epoch, self.started = modality, None, False

    def __init__(self, make, device='cpu'):
        self.make, self.pending, self.device = make, None, device

    def open(self, modality):
        t = _Plan._Ticket(modality)
        p = self.pending
        if p is not None and not p.started and p.modality != modality and p.epoch is None:
            p.epoch = t.epoch = [None]              # one shared cell, filled by whichever side asks for its batches first
            self.pending = None
        else:
            self.pending = t
        return t

    def batches(self, ticket):
        ticket.started = True
        if self.pending is ticket:
            self.pending = None
        if ticket.epoch is None:
            ticket.epoch = [None]
        if ticket.epoch[0] is None:
            # the epoch's index lists go to the device in ONE copy; a batch is a view of that tensor.  (One pageable host->device
            # copy per batch and modality was a sixth of a sweep job's host time, r2 profile, and a synchronisation point.)
            lists = [np.asarray(b, dtype=np.int64).reshape(-1) for b in self.make()]
            flat = torch.from_numpy(np.concatenate(lists) if lists else np.zeros(0, dtype=np.int64)).to(self.device)
            out, o = [], 0
            for b in lists:
                out.append(flat[o:o + len(b)])
                o += len(b)
            ticket.epoch[0] = out
        return ticket.epoch[0]


def build_loaders(data: PackedDataset, batch_size=100, training=True, random_state=789):
    """{'FFNN': loader, 'CNN': loader} as Kfold_CV_Multimodal.build_dataloader_forCV produces them (:537-546):
    training -> BalancePos_BatchSampler(batch_size); otherwise shuffled batches of 2 * batch_size drawn with
    torch.Generator().manual_seed(random_state + 30)."""
    if training:
        sampler = BalancePos_BatchSampler(data.labels_host, batch_size)
        plan, n_len = _Plan(lambda: iter(sampler), data.device), len(sampler)
    else:
        bs = 2 * batch_size
        gen = torch.Generator().manual_seed(random_state + 30)

        def make():
            perm = torch.randperm(len(data), generator=gen).tolist()
            return (perm[i:i + bs] for i in range(0, len(perm), bs))
        plan, n_len = _Plan(make, data.device), (len(data) + bs - 1) // bs
    return {'FFNN': DeviceLoader(data, 'FFNN', plan, n_len), 'CNN': DeviceLoader(data, 'CNN', plan, n_len)}
