"""Data-parallel training of one EmbraceNet over the GPUs of a node: one process per GPU,
`torch.distributed` (NCCL over NVLink) for the plumbing.

The global batch is partitioned by rows.  What couples the rows, and how it is kept equal to the single-GPU
large-batch step:
  * BatchNorm batch statistics      -> the engine hands its per-channel partial sums ([sum, sumsq] forward,
                                       [sum dz, sum dz*xhat] backward; 2*C doubles per conv layer) to an all-reduce
                                       between its stats and finalize kernels (SyncBN); n is the GLOBAL B*L
  * per-batch class weights of the loss -> computed from the GLOBAL positive count (labels are known to the host)
  * random draws                    -> Philox counters are keyed by GLOBAL row, so the partition does not change them
  * parameter gradients             -> all-reduce (sum) of the flat fp32 gradient arena, in two parts: the non-CNN slices
                                       start while the CNN backward still runs (engine phase hook), the CNN slice follows
No other collective exists on the path.
"""
import torch
import torch.distributed as dist


def shard_rows(global_batch, rank, world):
    """Contiguous, near-equal row ranges: rank r owns [lo, hi)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_step_metrics(records, group=None):
    """Each rank's EmbStepMetrics hold its share of the (globally normalised) loss and its confusion counts:
    the global record is their sum.  records: list of dicts; returns the merged list (same on every rank)."""
    if not records:
        return records
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    t = torch.tensor([[r['loss'], r['tp'], r['fp'], r['fn'], r['tn']] for r in records], dtype=torch.float64, device=dev)
    dist.all_reduce(t, group=group)
    t = t.cpu()
    return [dict(loss=float(a[0]), tp=int(a[1]), fp=int(a[2]), fn=int(a[3]), tn=int(a[4])) for a in t]


class DataParallel:
    """Wraps an Engine whose rows are this rank's shard of the global batch."""

    def __init__(self, engine, global_batch, rank=None, world=None, group=None, overlap=True, graph=False):
        self.engine, self.group = engine, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.global_batch = global_batch
        self.lo, self.hi = shard_rows(global_batch, self.rank, self.world)
        engine.set_shard(self.lo, global_batch)
        engine.set_allreduce(lambda t: dist.all_reduce(t, group=group))
        # Gradient all-reduce in two parts: everything outside the CNN stack is final before the CNN backward starts, so
        # those arena slices are reduced on a side stream while the CNN backward (most of the step) still runs.
        pc = 'CNN.' if engine.spec.kind == 'embracenet' else 'CNN_model.'
        self.cnn_lo, self.cnn_hi = engine.arena_range(pc)
        self.overlap = overlap and dist.get_backend(group) == 'nccl' and self.cnn_hi > self.cnn_lo
        self._pending = []
        # graph=True (opt-in, NCCL only): the whole step -- SyncBN all-reduces, both parts of the gradient all-reduce and the
        # optimizer kernel -- is captured into ONE CUDA graph by the engine (emb_set_graph(2)); the callbacks below then run
        # during capture only, on the capture stream, and every later step is a single graph launch.
        self.graph = bool(graph) and dist.get_backend(group) == 'nccl'
        if self.graph and int(graph) < 2:
            self.overlap = False            # graph=1: one all-reduce of the whole arena after the backward pass; graph=2: early slices
        if self.overlap:
            self.comm_stream = torch.cuda.Stream(device=engine.device)
        if self.overlap or self.graph:
            engine.set_phase_hook(self._grads_ready)
        if self.graph:
            engine.set_graph(True, collectives=True)

    def _grads_ready(self, phase):
        g = self.engine.grads
        if phase == 2:                      # graph mode: after the backward pass, before the optimizer kernel
            self._finish_gradients()
            return
        if not self.overlap:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.engine.device))
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            for lo, hi in ((0, self.cnn_lo), (self.cnn_hi, g.numel())):
                if hi > lo:
                    self._pending.append(dist.all_reduce(g[lo:hi], group=self.group, async_op=True))

    def close(self):
        """Graph mode: destroy the captured step graphs BEFORE the process group goes away -- NCCL keeps a communicator alive
        (and `destroy_process_group` blocks) for as long as a CUDA graph that captured its collectives exists."""
        if self.graph:
            torch.cuda.synchronize(self.engine.device)
            self.engine.set_graph(False)
            self.graph = False

    def broadcast_parameters(self, src=0):
        dist.broadcast(self.engine.params, src, group=self.group)
        dist.broadcast(self.engine.buffers, src, group=self.group)

    def _finish_gradients(self):
        eng = self.engine
        if self.overlap:
            dist.all_reduce(eng.grads[self.cnn_lo:self.cnn_hi], group=self.group)     # the CNN slice, after its backward
            for w in self._pending:
                w.wait()                                            # the current stream waits for the early slices
            self._pending = []
            if self.graph:                  # the side stream forked into the capture at phase 1: join it back
                ev = torch.cuda.Event()
                ev.record(self.comm_stream)
                torch.cuda.current_stream(eng.device).wait_event(ev)
        else:
            dist.all_reduce(eng.grads, group=self.group)           # the gradient all-reduce

    def train_step(self, x_local, bases_local, y_local, n_pos_global, cfg):
        eng = self.engine
        eng.set_global_positives(n_pos_global)
        if self.graph:
            eng.train_step(x_local, bases_local, y_local, cfg)    # one call = one graph launch (collectives and optimizer inside)
            return
        eng.train_step(x_local, bases_local, y_local, None)       # forward + loss + backward (SyncBN inside)
        self._finish_gradients()
        eng.opt_step(cfg)
