"""Data-parallel training of one EmbraceNet over the GPUs of a node: one process per GPU,
`torch.distributed` (NCCL over NVLink) for the plumbing.

The global batch is partitioned by rows.  What couples the rows, and how it is kept equal to the single-GPU
large-batch step:
  * BatchNorm batch statistics      -> the engine hands its per-channel partial sums ([sum, sumsq] forward,
                                       [sum dz, sum dz*xhat] backward; 2*C doubles per conv layer) to an all-reduce
                                       between its stats and finalize kernels (SyncBN); n is the GLOBAL B*L
  * per-batch class weights of the loss -> computed from the GLOBAL positive count (labels are known to the host)
  * random draws                    -> Philox counters are keyed by GLOBAL row, so the partition does not change them
  * parameter gradients             -> reduce-scatter of the flat fp32 gradient arena + optimizer on the owned slice +
                                       all-gather of the updated parameters, ONE kernel over peer memory (comm='peer');
                                       or an NCCL all-reduce of the arena in two parts (comm='nccl', the round-1 path)
No other exchange exists on the path.
"""
import ctypes as C

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


def _ptr_table(ptrs):
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])


def attach_peers(engine, rank, world, comm_ptrs, params_ptrs, grads_ptrs):
    """emb_dp_attach with pointer tables that are valid in this process (entry [rank] = the engine's own buffers)."""
    from . import _native as N
    N.check(engine.lib.emb_dp_attach(engine._h, rank, world, _ptr_table(comm_ptrs), _ptr_table(params_ptrs), _ptr_table(grads_ptrs)))


def new_comm_block(device):
    """A zeroed, 128-byte aligned communication block (emb_dp_comm_bytes) as a uint8 tensor view."""
    from . import _native as N
    n = int(N.lib().emb_dp_comm_bytes())
    raw = torch.zeros(n + 128, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 128
    return raw, raw[off:off + n]


def attach_local(engines, global_batch):
    """Data parallelism between engines that share ONE process (same or different devices): shards are assigned in list
    order; pointers are exchanged directly.  Returns the comm blocks (keep them alive).  Used by the single-GPU DP test:
    each engine must then be driven on its own CUDA stream, and a step's calls issued for every engine before any
    synchronisation (an exchange kernel waits for its peers on the device)."""
    world = len(engines)
    blocks = [new_comm_block(e.device) for e in engines]
    comm = [b[1].data_ptr() for b in blocks]
    params = [e.params.data_ptr() for e in engines]
    grads = [e.grads.data_ptr() for e in engines]
    for r, e in enumerate(engines):
        lo, hi = shard_rows(global_batch, r, world)
        e.set_shard(lo, global_batch)
        attach_peers(e, r, world, comm, params, grads)
    return blocks


class DataParallel:
    """Wraps an Engine whose rows are this rank's shard of the global batch.

    comm='peer' (default): every exchange of the step is a kernel of the library over NVLink peer memory (csrc/dp_peer.cuh);
    torch.distributed only carries the CUDA IPC handles once, at construction.  The step is then captured in a CUDA graph
    like the single-GPU step.  comm='nccl': the round-1 host-driven path (SyncBN sums and the gradient arena all-reduced by
    NCCL from callbacks), kept as the A/B reference and as the fall-back when peer mapping is unavailable."""

    def __init__(self, engine, global_batch, rank=None, world=None, group=None, overlap=True, graph=True, comm='peer'):
        self.engine, self.group = engine, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.global_batch = global_batch
        self.lo, self.hi = shard_rows(global_batch, self.rank, self.world)
        engine.set_shard(self.lo, global_batch)
        self.comm = comm
        self._pending = []
        self.graph = False
        if comm == 'peer' and dist.get_backend(group) == 'nccl' and self.world > 1:
            ok = self._attach_ipc()
            flag = torch.tensor([1 if ok else 0], device=engine.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                if graph:
                    engine.set_graph(True)
                    self.graph = True
                return
            engine.lib.emb_dp_detach(engine._h)
            import sys
            print('embrace_b200.dp: peer-memory mapping unavailable, falling back to NCCL collectives', file=sys.stderr)
        self.comm = 'nccl'
        self._init_nccl(overlap, graph if comm == 'nccl' else False)

    # ---- peer-memory mode -------------------------------------------------------------------------------------------
    def _attach_ipc(self):
        from . import _native as N
        eng = self.engine
        try:
            self._comm_raw, self._comm = new_comm_block(eng.device)
            torch.cuda.synchronize(eng.device)
            mine = []
            for t in (self._comm, eng.params, eng.grads):
                h, off = (C.c_char * 64)(), C.c_int64()
                N.check(eng.lib.emb_ipc_export(C.c_void_p(t.data_ptr()), h, C.byref(off)))
                mine.append((bytes(h), int(off.value)))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=self.group)
            tables = [[], [], []]
            with torch.cuda.device(eng.device):
                for q, entry in enumerate(everyone):
                    for k, (h, off) in enumerate(entry):
                        if q == self.rank:
                            tables[k].append((self._comm, eng.params, eng.grads)[k].data_ptr())
                        else:
                            out = C.c_void_p()
                            N.check(eng.lib.emb_ipc_open(C.c_char_p(h), off, C.byref(out)))
                            tables[k].append(out.value)
                attach_peers(eng, self.rank, self.world, *tables)
            return True
        except Exception as ex:                      # noqa: BLE001 -- any failure here means "use the NCCL path"
            import sys
            print(f'embrace_b200.dp: peer attach failed on rank {self.rank}: {ex}', file=sys.stderr)
            return False

    # ---- NCCL mode (round-1 path) -----------------------------------------------------------------------------------
    def _init_nccl(self, overlap, graph):
        engine, group = self.engine, self.group
        engine.set_allreduce(lambda t: dist.all_reduce(t, group=group))
        # Gradient all-reduce in two parts: everything outside the CNN stack is final before the CNN backward starts, so
        # those arena slices are reduced on a side stream while the CNN backward (most of the step) still runs.
        pc = 'CNN.' if engine.spec.kind == 'embracenet' else 'CNN_model.'
        self.cnn_lo, self.cnn_hi = engine.arena_range(pc)
        self.overlap = overlap and dist.get_backend(group) == 'nccl' and self.cnn_hi > self.cnn_lo
        # graph (opt-in, NCCL only): the whole step -- SyncBN all-reduces, both parts of the gradient all-reduce and the
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
        """Destroy the captured step graphs (NCCL mode: BEFORE the process group goes away -- NCCL keeps a communicator alive,
        and `destroy_process_group` blocks, for as long as a CUDA graph that captured its collectives exists) and, in peer
        mode, make sure no rank unmaps memory a peer's kernel may still touch."""
        torch.cuda.synchronize(self.engine.device)
        if self.comm == 'peer':
            dist.barrier(group=self.group)
            self.engine.lib.emb_dp_detach(self.engine._h)
        if self.graph:
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
        """One data-parallel train step on this rank's rows.  n_pos_global (the positive count of the GLOBAL batch) is only
        needed by the NCCL mode; the peer mode exchanges the counts on the device and accepts None."""
        eng = self.engine
        if self.comm == 'peer':
            eng.train_step(x_local, bases_local, y_local, cfg)    # exchanges and the sharded optimizer are inside (one graph launch)
            return
        eng.set_global_positives(n_pos_global)
        if self.graph:
            eng.train_step(x_local, bases_local, y_local, cfg)    # one call = one graph launch (collectives and optimizer inside)
            return
        eng.train_step(x_local, bases_local, y_local, None)       # forward + loss + backward (SyncBN inside)
        self._finish_gradients()
        eng.opt_step(cfg)

    def train_step_host(self, x_host, bases_host, y_host, cfg):
        """Pinned host buffers in, the PREVIOUS step's metrics record out (peer mode): the software-pipelined host entry of the
        engine, with the data-parallel exchanges inside the step."""
        assert self.comm == 'peer', 'the pipelined host entry needs the peer-memory mode'
        return self.engine.train_step_host_pipelined(x_host, bases_host, y_host, cfg)

    def allreduce_grads(self):
        """Peer mode, tests: leave the summed gradient in every rank's gradient arena."""
        from . import _native as N
        N.check(self.engine.lib.emb_dp_allreduce_grads(self.engine._h, self.engine.stream))
