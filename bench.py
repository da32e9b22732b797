#!/usr/bin/env python
"""bench.py -- EmbraceNet train samples/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload L8192|S256|infer] [--arch ..] [--batch ..]

Workloads (BASELINE.json configs; the default is the one the metric is quoted on):
  L8192 (default)  configs[2]: full EmbraceNet training, arch "L" (largest point of the search space), synthetic HEPG2-promoter
                   shape (F=562, 256-bp bases), GLOBAL batch 8192 split over the N ranks (strong scaling), Adam + coupled L2
  S256             configs[0]: arch "S" (real best trial of the A549 notebook), F=48, batch 256 -- the small-batch step the
                   hyper-parameter sweeps (configs[3]) actually run
  infer            configs[4]: eval-mode scoring of synthetic regions, arch S, availability mix 80/10/10, row shards per rank
A "step" is one pass of the hot path over one batch.

  value     device-resident inputs, emb_train_step (forward + loss + backward + optimizer; one CUDA-graph launch; under data
            parallelism the SyncBN / loss-weight / gradient exchanges are kernels inside that graph), CUDA-event timed;
            the K-step block is repeated until >= --min-seconds of device time have been measured (mean per step reported)
  e2e       the same through the pipelined host entry: pinned HOST buffers in, EmbStepMetrics back, copies inside the timed region
  roofline  the GEMM kernel class (Conv1d implicit GEMMs, docking, Linear; fwd/dgrad/wgrad), timed per launch with CUDA
            event pairs on the launching stream inside timed steps
  cpu_baseline  oracle/torch_port.py (the reference's own PyTorch-CPU library calls, fp64) on a bounded sample

--impl reference times that CPU port alone, with all host threads, on the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'embracenet_train_samples_per_sec'
UNIT = 'samples/s'

WORKLOADS = {
    'L8192': dict(arch='L', batch=8192, config='BASELINE configs[2]'),
    'S256': dict(arch='S', batch=256, config='BASELINE configs[0]'),
    'infer': dict(arch='S', batch=65536, config='BASELINE configs[4]'),
}


def oracle_spec(name):
    """Spec dict of the CPU oracle / port (test infrastructure; used by the cpu_baseline and --impl reference legs only)."""
    from tests.golden.cases import ARCH_L, ARCH_S, ARCH_M, ARCH_W
    return {'L': ARCH_L, 'S': ARCH_S, 'M': ARCH_M, 'W': ARCH_W}[name]


def cpu_port_sample(arch_name, batch, budget_s=14.0, max_steps=64):
    """Time the PyTorch-CPU port of the reference on a bounded sample: one warm-up, one probe step, then as many steps
    as fit `budget_s` seconds."""
    import torch
    from oracle import torch_port as TP
    spec = oracle_spec(arch_name)
    torch.set_num_threads(os.cpu_count() or 1)
    probe = TP.time_train(spec, batch, 1, 1, seed=789)
    steps = int(max(2, min(max_steps, budget_s / max(probe['seconds'], 1e-3))))
    r = TP.time_train(spec, batch, steps, 0, seed=789)
    r['steps'] = steps
    return r


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def workload_text(args, F):
    w = WORKLOADS[args.workload]
    if args.workload == 'infer':
        return (f'EmbraceNet arch {args.arch} eval-mode scoring, {args.batch} synthetic regions per step and rank, availability 80/10/10 '
                f'(both / epigenomic only / sequence only), F={F}, 256-bp bases ({w["config"]})')
    return (f'EmbraceNet arch {args.arch} full train step (fwd+loss+bwd+Adam), global batch {args.batch}, F={F}, 256-bp bases ({w["config"]})')


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (PyTorch fp64 library calls, oracle/torch_port.py), all host threads.
    Each step is one train step at the SAME global batch as the measured arm when K steps of it fit the time budget
    (--ref-budget seconds); otherwise the batch is halved until they do, and the config says so (same_config false)."""
    if rank != 0:
        return
    import torch
    from oracle import torch_port as TP
    spec = oracle_spec(args.arch)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    if args.workload == 'infer':
        return run_reference_infer(args, spec, threads)
    B = args.ref_batch or args.batch
    probe = None
    while True:
        probe = TP.time_train(spec, B, 1, 1, seed=789)              # one warm-up step + one timed step
        if args.ref_batch or probe['seconds'] * args.steps <= args.ref_budget or B <= 256:
            break
        B //= 2
    r = TP.time_train(spec, B, args.steps, 0, seed=789)            # the probe above was the warm-up (the CPU path has no lazy state left after it)
    val = r['samples_per_s']
    sample = (f'{args.steps} train steps of batch {B} after 2 warm-up (arch {args.arch}, fp64, torch {torch.__version__} CPU, '
              f'{r["seconds"]:.1f} s)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * r['seconds'] / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_text(args, spec['F']), 'arch': args.arch, 'global_batch': args.batch, 'reference_batch': B,
                   'same_config': B == args.batch, 'in_features': spec['F'], 'optimizer': 'adam+L2'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': r['threads'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_reference_infer(args, spec, threads):
    """CPU arm of the inference workload: the port's eval forward + softmax on bounded batches (batch 4096 per step)."""
    import numpy as np
    import torch
    from oracle import embracenet_oracle as O
    from oracle import torch_port as TP
    from tests.golden.cases import make_inputs
    B = args.ref_batch or 4096
    P = O.init_params(spec, 789)
    T = TP.params_to_torch(P, requires_grad=False)
    x, bases, _ = make_inputs(spec, B, 790)
    x1, x2 = torch.from_numpy(x), torch.from_numpy(O.onehot_from_bases(bases))
    av = np.ones((B, 2), dtype=np.float32)
    u = np.random.RandomState(5).random_sample(B)
    av[(u >= 0.8) & (u < 0.9), 1] = 0
    av[u >= 0.9, 0] = 0

    def step():
        with torch.no_grad():
            lg, _ = TP.forward(spec, T, x1, x2, None, training=False, availabilities=av)
            return torch.softmax(lg, dim=1)[:, 1]
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    line = {'impl': 'reference', 'metric': 'embracenet_infer_regions_per_sec', 'value': val, 'unit': 'regions/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_text(args, spec['F']), 'arch': args.arch, 'reference_batch': B, 'same_config': False},
            'cpu_baseline': {'value': val, 'unit': 'regions/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{args.steps} eval forwards of batch {B} (fp64 PyTorch CPU port, {dt:.1f} s)'},
            'e2e': {'value': val, 'unit': 'regions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='L8192', choices=list(WORKLOADS))
    ap.add_argument('--arch', default=None)
    ap.add_argument('--batch', type=int, default=None, help='GLOBAL batch (train) / regions per step and rank (infer)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--tensor-core', type=int, default=1)
    ap.add_argument('--ref-batch', type=int, default=0, help='force the batch of the CPU arm (0: the global batch, halved until it fits --ref-budget)')
    ap.add_argument('--ref-budget', type=float, default=0.0,
                    help='seconds the timed steps of --impl reference may take (default: 300 at N = 1 -- K = 20 steps of the full batch 8192 fit -- and 120 for N > 1)')
    ap.add_argument('--cpu-baseline', type=int, default=1)
    ap.add_argument('--min-seconds', type=float, default=2.0, help='repeat the K-step timed block until this much device time is measured')
    ap.add_argument('--graph', type=int, default=1, help='replay the train step as one CUDA graph')
    ap.add_argument('--dp-comm', default=os.environ.get('EMB_DP_COMM', 'peer'), choices=['peer', 'nccl'],
                    help='N > 1: peer = exchanges as library kernels over NVLink peer memory (default); nccl = round-1 host-driven collectives')
    ap.add_argument('--dp-graph', type=int, default=int(os.environ.get('EMB_DP_GRAPH', '0')), help='nccl mode only: capture the NCCL collectives in the step graph')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    args.arch = args.arch or w['arch']
    args.batch = args.batch or w['batch']
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 1)

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        if args.ref_budget <= 0:
            args.ref_budget = 300.0 if world == 1 else 120.0
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import embrace_b200
    from embrace_b200 import presets

    spec = presets.arch(args.arch)            # product-side spec: the measured arm imports nothing from oracle/ or tests/
    F = spec.in_features

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # stdout carries exactly one JSON line: NCCL's own log (whatever NCCL_DEBUG level the caller chose) goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_blocks(fn, steps, min_seconds, finish=None):
        """K-step blocks, each bracketed by CUDA events (barrier + synchronize around the whole series), repeated until
        min_seconds of device time; returns (mean ms per step, blocks, [ms per block])."""
        per_block, total, blocks = [], 0.0, 0
        while True:
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(blocks * steps + i)
            if finish:
                finish()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            per_block.append(ms)
            total += ms
            blocks += 1
            if total >= min_seconds * 1e3 or blocks >= 2000:
                break
        return total / (blocks * steps), blocks, per_block

    peaks = {}
    pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))

    if args.workload == 'infer':
        return run_infer(args, spec, dev, rank, world, local_rank, barrier, timed_blocks, peaks)

    if args.batch % world:
        raise SystemExit('global batch must divide by the number of ranks')
    B = args.batch // world
    NBUF = 4

    eng = embrace_b200.Engine(spec, max_batch=B, precision=args.precision, device=dev, seed=789,
                              tensor_core=bool(args.tensor_core) and args.precision == 'bf16')
    eng.init_random(789)                                   # same random-init weights on every rank
    cfg = eng.opt_config('adam', lr=4.1e-5, weight_decay=7.6e-4)
    glob = presets.synthetic_batches(spec, args.batch, NBUF, seed=789)   # every rank builds the global batches, keeps its rows
    lo = rank * B
    host = [(torch.from_numpy(x[lo:lo + B]).pin_memory(), torch.from_numpy(b[lo:lo + B]).pin_memory(),
             torch.from_numpy(y[lo:lo + B]).pin_memory()) for x, b, y in glob]
    devb = [(x.to(dev), b.to(dev), y.to(dev)) for x, b, y in host]
    npos = [int(y.sum()) for _, _, y in glob]
    dp = None
    if world > 1:
        from embrace_b200.dp import DataParallel
        dp = DataParallel(eng, args.batch, rank, world, comm=args.dp_comm, graph=bool(args.graph) if args.dp_comm == 'peer' else args.dp_graph)
        assert (dp.lo, dp.hi) == (lo, lo + B)
    elif args.graph:
        eng.set_graph(True)
    peer = dp is not None and dp.comm == 'peer'

    def step_device(i):
        x, b, y = devb[i % NBUF]
        if dp is not None:
            dp.train_step(x, b, y, npos[i % NBUF], cfg)      # SyncBN sums, global loss weights, gradient reduction + sharded optimizer
        else:
            eng.train_step(x, b, y, cfg)

    def step_host(i):
        x, b, y = host[i % NBUF]
        if dp is not None and not peer:
            xd, bd, yd = x.to(dev, non_blocking=True), b.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            eng.metrics_reset()
            dp.train_step(xd, bd, yd, npos[i % NBUF], cfg)
            return eng.metrics_read(1)[0]['loss']
        return eng.train_step_host_pipelined(x, b, y, cfg)      # copy of batch i overlaps step i-1; returns step i-1's record

    for i in range(args.warmup):
        step_device(i)
    barrier()
    # pass 1 (eager, per-launch CUDA event pairs around the GEMM class): the roofline numbers
    eng.profile_gemm(True)
    ms_prof, _, _ = timed_blocks(step_device, args.steps, 0.0)
    prof = eng.profile_read()
    eng.profile_gemm(False)
    for i in range(2):
        step_device(i)                                       # re-enter the graph path
    # pass 2: `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    ms_step, blocks, per_block = timed_blocks(step_device, args.steps, args.min_seconds)
    launches = (eng.launch_count - l0) / blocks
    clocks = sampler.stop() if rank == 0 else None

    pipelined = dp is None or peer
    for i in range(2):
        step_host(i)
    if pipelined:
        eng.flush_host()
    ms_e2e, blocks_e2e, _ = timed_blocks(step_host, args.steps, args.min_seconds, finish=eng.flush_host if pipelined else None)
    final = eng.metrics_read(4)
    if dp is not None:
        from embrace_b200.dp import merge_step_metrics
        final = final[-1:] or [dict(loss=0.0, tp=0, fp=0, fn=0, tn=0)]      # the same shape on every rank, whatever was recorded
        final = merge_step_metrics(final)        # every rank holds its share of the globally normalised loss: report the sum
        dp.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_sus, peak_burst = peaks.get('bf16_tflops_sustained', 1400.0), peaks.get('bf16_tflops', 1600.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained; frac_burst uses bf16_tflops)' if peaks else 'fallback (B200_PROFILING.md)'
    achieved_tf = prof['flops'] / (prof['ms'] * 1e-3) / 1e12 if prof['ms'] > 0 else 0.0
    # DRAM traffic of the GEMM kernel class per launch, from the committed ncu capture of this same command
    # (profiles/summarize_launches.py); only valid for the default workload
    traffic = None
    for name in ('r02_gemm_traffic.json', 'r01_gemm_traffic.json'):
        tr_path = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(tr_path) and args.workload == 'L8192' and args.arch == 'L' and args.batch == 8192 and world == 1:
            traffic = json.load(open(tr_path)).get('gemm_dram_bytes_per_launch')
            break

    value = args.batch / (ms_step * 1e-3)
    e2e = args.batch / (ms_e2e * 1e-3)
    gemm_ms = prof['ms'] / args.steps
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_text(args, F), 'arch': args.arch, 'global_batch': args.batch, 'per_gpu_batch': B, 'in_features': F,
                   'optimizer': 'adam+L2', 'precision': args.precision, 'tensor_core': eng.tensor_core,
                   'cuda_graph': bool(args.graph) if (world == 1 or peer) else bool(args.dp_graph),
                   'parallelism': (f'dp{world} ({dp.comm}: ' + ('SyncBN / loss-weight / gradient exchanges as kernels over NVLink peer memory, sharded optimizer)'
                                                                if peer else 'NCCL all-reduces driven from the host)')) if world > 1 else 'single',
                   'l2': f'per-step working set (activations+gradients, ~{eng.ws_bytes / 1e9:.2f} GB) vs 126 MB L2; '
                         f'inputs rotate over {NBUF} resident batches',
                   'timed': f'{blocks} blocks of {args.steps} steps (>= {args.min_seconds} s of device time), mean; block min/max '
                            f'{min(per_block) / args.steps:.4f}/{max(per_block) / args.steps:.4f} ms per step',
                   'train_flops_per_sample': 3 * presets.fwd_flops_per_sample(spec)},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': UNIT, 'ms_per_step': ms_e2e, 'blocks': blocks_e2e,
                'h2d_bytes_per_step': int(args.batch * (F * 4 + 256 + 4)), 'd2h_bytes_per_step': 20 * world},
        'gpu_launches': int(round(launches * blocks)),
        'gpu_launches_per_step': launches / args.steps,
        'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_sus, 'unit': 'TFLOP/s',
                     'frac': achieved_tf / peak_sus, 'frac_burst': achieved_tf / peak_burst, 'peak_burst': peak_burst,
                     'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu, mean over the class)',
                     'algorithmic_flops_per_launch': prof['flops'] / max(prof['launches'], 1),
                     'kernel': 'GEMM class: tc_gemm_kernel + tc_conv_reuse_kernel (conv implicit GEMMs, docking with the embracement epilogue, linear; fwd/dgrad/wgrad)',
                     'kernel_ms_per_step': gemm_ms, 'kernel_launches_per_step': prof['launches'] / args.steps,
                     'kernel_share_of_step': gemm_ms / max(ms_step, 1e-9), 'eager_profiled_ms_per_step': ms_prof,
                     'peak_source': peak_src},
        'final_loss': final[-1]['loss'] if final else None,
    }
    if world == 1 and args.cpu_baseline:
        cb = min(args.batch, 2048)
        r = cpu_port_sample(args.arch, cb)
        line['cpu_baseline'] = {'value': r['samples_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
                                'sample': f'{r["steps"]} train steps of batch {cb} after 1 warm-up (arch {args.arch}, fp64 PyTorch CPU '
                                          f'port of the reference, {r["seconds"]:.1f} s of CPU work)'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_infer(args, spec, dev, rank, world, local_rank, barrier, timed_blocks, peaks):
    """configs[4]: every rank scores its own row shard (no collective on the data path); a step = args.batch regions per rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import embrace_b200
    from embrace_b200 import presets
    from embrace_b200.infer import synthetic_availabilities
    F, B = spec.in_features, args.batch
    eng = embrace_b200.Engine(spec, max_batch=B, precision=args.precision, device=dev, seed=789,
                              tensor_core=bool(args.tensor_core) and args.precision == 'bf16')
    eng.init_random(789)
    NBUF = 3
    rs = np.random.RandomState(1000 + rank)
    host = []
    for k in range(NBUF):
        x = torch.from_numpy(rs.random_sample((B, F)).astype(np.float32)).pin_memory()
        codes = torch.from_numpy(rs.randint(0, 4, size=(B, 256)).astype(np.uint8)).pin_memory()
        av = torch.from_numpy(synthetic_availabilities(B, seed=rank * 10 + k)).pin_memory()
        host.append((x, codes, av))
    devb = [(x.to(dev), c.to(dev), a.to(dev)) for x, c, a in host]
    out = torch.empty(B, dtype=torch.float32, device=dev)

    def step_device(i):
        x, c, a = devb[i % NBUF]
        eng.infer(x, c, a, out)

    out_host = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]

    def step_host(i):
        # the scoring loop a user runs over host batches: batch i uploads while batch i-1 computes; every batch's inputs cross the
        # host link and every batch's scores come back to host memory inside the timed region (flush at the end of each block)
        x, c, a = host[i % NBUF]
        return eng.predict_host_pipelined(x, c, a, out_host[i & 1])

    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    ms_step, blocks, per_block = timed_blocks(step_device, args.steps, args.min_seconds)
    launches = (eng.launch_count - l0) / blocks
    clocks = sampler.stop() if rank == 0 else None
    step_host(0)
    eng.predict_host_flush()
    ms_e2e, blocks_e2e, _ = timed_blocks(step_host, args.steps, args.min_seconds, finish=eng.predict_host_flush)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    regions = B * world
    value, e2e = regions / (ms_step * 1e-3), regions / (ms_e2e * 1e-3)
    flops = presets.fwd_flops_per_sample(spec)
    peak_sus = peaks.get('bf16_tflops_sustained', 1400.0)
    achieved_tf = flops * B / (ms_step * 1e-3) / 1e12
    line = {
        'metric': 'embracenet_infer_regions_per_sec', 'value': value, 'unit': 'regions/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_text(args, F), 'arch': args.arch, 'regions_per_step_per_gpu': B, 'in_features': F,
                   'precision': args.precision, 'parallelism': f'row shards x{world}, no collective',
                   'timed': f'{blocks} blocks of {args.steps} steps, mean; block min/max {min(per_block) / args.steps:.4f}/{max(per_block) / args.steps:.4f} ms',
                   'fwd_flops_per_region': flops},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': 'regions/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': int(B * (F * 4 + 256 + 8)) * world,
                'd2h_bytes_per_step': 4 * B * world},
        'gpu_launches': int(round(launches * blocks)), 'gpu_launches_per_step': launches / args.steps,
        'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_sus, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_sus,
                     'traffic': None, 'kernel': 'whole eval forward (per GPU): dense-equivalent forward FLOPs / step time',
                     'algorithmic_bytes_per_region': F * 4 + 256 + 8 + 4},
    }
    if world == 1 and args.cpu_baseline:
        import io
        import contextlib
        buf = io.StringIO()
        a2 = argparse.Namespace(**vars(args))
        a2.steps, a2.warmup, a2.ref_batch = 3, 1, 4096
        with contextlib.redirect_stdout(buf):
            run_reference_infer(a2, oracle_spec(args.arch), os.cpu_count() or 1)
        line['cpu_baseline'] = json.loads(buf.getvalue())['cpu_baseline']
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
