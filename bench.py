#!/usr/bin/env python
"""bench.py -- EmbraceNet train samples/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--arch L|S|M|W] [--batch B] [--precision bf16|fp32]

Workload (BASELINE.json configs[2]): full EmbraceNet training, arch "L" (largest point of the
search space), synthetic HEPG2-promoter shape (F=562, 256-bp bases), GLOBAL batch 8192 split over the N ranks
(strong scaling), Adam with coupled L2.  A "step" is one pass of the train-step hot path over one batch.

  value   device-resident inputs, emb_train_step (forward + loss + backward + optimizer), CUDA-event timed
  e2e     the same through emb_train_step_host: pinned HOST buffers in, EmbStepMetrics back, copies inside
          the timed region
  roofline  the GEMM kernel class (Conv1d implicit GEMMs, docking, Linear; fwd/dgrad/wgrad), timed per launch
          with CUDA event pairs inside the timed steps
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


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (PyTorch fp64 library calls, oracle/torch_port.py), all host threads.
    Each step is a bounded sample of the workload: one train step of batch --ref-batch."""
    if rank != 0:
        return
    import torch
    from oracle import torch_port as TP
    spec = oracle_spec(args.arch)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = args.ref_batch
    r = TP.time_train(spec, B, args.steps, args.warmup, seed=789)
    val = r['samples_per_s']
    sample = f'{args.steps} train steps of batch {B} after {args.warmup} warm-up (arch {args.arch}, fp64, torch {torch.__version__} CPU, {r["seconds"]:.1f} s)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * r['seconds'] / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'EmbraceNet arch {args.arch} full train step (fwd+loss+bwd+Adam), global batch {args.batch}, '
                               f'F={spec["F"]}, 256-bp bases (BASELINE configs[2]); bounded CPU sample: batch {B} per step',
                   'arch': args.arch, 'global_batch': args.batch, 'in_features': spec['F'], 'optimizer': 'adam+L2'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': r['threads'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--arch', default='L')
    ap.add_argument('--batch', type=int, default=8192, help='GLOBAL batch')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--tensor-core', type=int, default=1)
    ap.add_argument('--ref-batch', type=int, default=256, help='batch of the bounded CPU sample')
    ap.add_argument('--cpu-baseline', type=int, default=1)
    ap.add_argument('--graph', type=int, default=1, help='replay the train step as one CUDA graph (single GPU)')
    ap.add_argument('--dp-graph', type=int, default=int(os.environ.get('EMB_DP_GRAPH', '0')),
                    help='N > 1: capture the data-parallel step with its NCCL collectives into one CUDA graph (1: one gradient '
                         'all-reduce after the backward pass, 2: early slices on a side stream); 0 = host-driven collectives')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 1)

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
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
        os.environ['NCCL_DEBUG'] = os.environ.get('EMB_NCCL_DEBUG', 'WARN')     # stdout carries exactly one JSON line
        dist.init_process_group('nccl', device_id=dev)
    if args.batch % world:
        raise SystemExit('global batch must divide by the number of ranks')
    B = args.batch // world
    NBUF = 4

    eng = embrace_b200.Engine(spec, max_batch=B, precision=args.precision, device=dev, seed=789,
                              tensor_core=bool(args.tensor_core) and args.precision == 'bf16')
    eng.init_random(789)                                   # same random-init weights on every rank
    if world == 1 and args.graph:
        eng.set_graph(True)
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
        dp = DataParallel(eng, args.batch, rank, world, graph=bool(args.dp_graph))
        assert (dp.lo, dp.hi) == (lo, lo + B)

    def step_device(i):
        x, b, y = devb[i % NBUF]
        if dp is not None:
            dp.train_step(x, b, y, npos[i % NBUF], cfg)      # SyncBN partial sums + ONE gradient all-reduce per step
        else:
            eng.train_step(x, b, y, cfg)

    def step_host(i):
        x, b, y = host[i % NBUF]
        if dp is not None:
            xd, bd, yd = x.to(dev, non_blocking=True), b.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            eng.metrics_reset()
            dp.train_step(xd, bd, yd, npos[i % NBUF], cfg)
            return eng.metrics_read(1)[0]['loss']
        return eng.train_step_host_pipelined(x, b, y, cfg)      # copy of batch i overlaps step i-1; returns step i-1's record

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    eng.profile_gemm(True)
    ms = timed(step_device, args.steps)
    prof = eng.profile_read()
    eng.profile_gemm(False)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    # second pass without the per-launch events (they serialise nothing, but keep `value` free of them)
    ms_clean = timed(step_device, args.steps)
    ms = min(ms, ms_clean)

    for i in range(2):
        step_host(i)
    if world == 1:
        eng.flush_host()

    def e2e_loop(i):
        step_host(i)
        if world == 1 and i == args.steps - 1:
            eng.flush_host()                 # the last step's record is read inside the timed region too
    ms_e2e = timed(e2e_loop, args.steps)
    final = eng.metrics_read(4)
    if dp is not None:
        from embrace_b200.dp import merge_step_metrics
        final = final[-1:] or [dict(loss=0.0, tp=0, fp=0, fn=0, tn=0)]      # the same shape on every rank, whatever was recorded
        final = merge_step_metrics(final)        # every rank holds its share of the globally normalised loss: report the sum
        dp.close()                               # graphs that captured NCCL collectives must be destroyed before their communicator

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md)'
    achieved_tf = prof['flops'] / (prof['ms'] * 1e-3) / 1e12 if prof['ms'] > 0 else 0.0
    # DRAM traffic of the GEMM kernel class per launch, from the committed ncu capture of this same command
    # (profiles/summarize_launches.py -> profiles/r01_gemm_traffic.json); only valid for the default workload
    traffic = None
    tr_path = os.path.join(ROOT, 'profiles', 'r01_gemm_traffic.json')
    if os.path.exists(tr_path) and args.arch == 'L' and args.batch == 8192 and world == 1:
        traffic = json.load(open(tr_path)).get('gemm_dram_bytes_per_launch')

    value = args.batch * args.steps / (ms * 1e-3)
    e2e = args.batch * args.steps / (ms_e2e * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': f'EmbraceNet arch {args.arch} full train step (fwd+loss+bwd+Adam), global batch {args.batch}, '
                               f'F={F}, 256-bp bases (BASELINE configs[2])',
                   'arch': args.arch, 'global_batch': args.batch, 'per_gpu_batch': B, 'in_features': F,
                   'optimizer': 'adam+L2', 'precision': args.precision, 'tensor_core': eng.tensor_core, 'cuda_graph': bool(args.graph if world == 1 else args.dp_graph),
                   'parallelism': f'dp{world}' if world > 1 else 'single',
                   'l2': f'per-step working set (activations+gradients, ~{eng.ws_bytes / 1e9:.1f} GB) >> 126 MB L2; '
                         f'inputs rotate over {NBUF} resident batches',
                   'train_flops_per_sample': 3 * presets.fwd_flops_per_sample(spec)},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': int(args.batch * (F * 4 + 256 + 4)), 'd2h_bytes_per_step': 20 * world},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                     'frac': achieved_tf / peak_tf, 'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu, mean over the class)',
                     'algorithmic_flops_per_launch': prof['flops'] / max(prof['launches'], 1),
                     'kernel': 'GEMM class: tc_gemm_kernel + tc_conv_reuse_kernel (conv implicit GEMMs, docking with the embracement epilogue, linear; fwd/dgrad/wgrad)',
                     'kernel_ms_per_step': prof['ms'] / args.steps, 'kernel_launches_per_step': prof['launches'] / args.steps,
                     'kernel_share_of_step': prof['ms'] / max(ms, 1e-9) if ms_clean >= ms else prof['ms'] / max(ms_clean, 1e-9),
                     'peak_source': peak_src},
        'final_loss': final[-1]['loss'] if final else None,
    }
    if world == 1 and args.cpu_baseline:
        r = cpu_port_sample(args.arch, args.ref_batch)
        line['cpu_baseline'] = {'value': r['samples_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
                                'sample': f'{r["steps"]} train steps of batch {args.ref_batch} after 1 warm-up (arch {args.arch}, fp64 PyTorch CPU '
                                          f'port of the reference, {r["seconds"]:.1f} s of CPU work)'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
