"""Per-kernel roofline table of one train step (arch L, batch 8192) from an ncu launch list:
python profiles/roofline_table.py profiles/r01_launches_archL_b8192.csv > profiles/r01_roofline_per_kernel.txt

GEMM-class launches are matched, in launch order, with their algorithmic FLOPs (2 M N K; conv: 2 B L Cin Cout taps) and rated
against the measured sustained bf16 peak; every other kernel is rated by its measured DRAM traffic / time against the measured
HBM copy bandwidth (both from MEASURED_PEAKS.json; per-launch times under ncu are cold-cache and serialised)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_launches import load, GEMM     # noqa: E402

B = 8192
LIN = lambda m, n, k: 2.0 * m * n * k                                  # noqa: E731
CONV = lambda L, ci, co: 2.0 * B * L * ci * co * 15                    # noqa: E731
# forward issue order (engine.cu forward_impl): the FFNN chain and docking_0 are issued first (on the side stream when the step forks),
# then the CNN chain, then docking_1 after the join
FWD = [('ffnn0 fwd', LIN(B, 256, 562)), ('ffnn1 fwd', LIN(B, 128, 256)), ('ffnn2 fwd', LIN(B, 64, 128)), ('ffnn3 fwd', LIN(B, 32, 64)),
       ('docking_0 fwd', LIN(B, 1024, 32)),
       ('conv1 fwd', CONV(124, 64, 96)), ('conv2 fwd', CONV(58, 96, 256)), ('conv3 fwd', CONV(25, 256, 512)),
       ('docking_1 fwd + embrace select', LIN(B, 1024, 4096)),
       ('post0 fwd', LIN(B, 512, 1024)), ('post1 fwd', LIN(B, 256, 512))]
# backward launch order (engine.cu backward_impl / cnn_backward_t): for every layer the WEIGHT gradient is launched first, then the
# data gradient (round 1 printed these two labels swapped); the 2-logit head runs on its own dot-product kernels, outside this class
_HEAD = [('post1 wgrad', LIN(B, 256, 512)), ('post1 dgrad', LIN(B, 256, 512)), ('post0 wgrad', LIN(B, 512, 1024)), ('post0 dgrad + embrace bwd', LIN(B, 512, 1024)),
         ('docking_0 wgrad', LIN(B, 1024, 32)), ('docking_0 dgrad', LIN(B, 1024, 32))]
_DOCK1 = [('docking_1 wgrad', LIN(B, 1024, 4096)), ('docking_1 dgrad', LIN(B, 1024, 4096))]
_FFNN = [('ffnn3 wgrad', LIN(B, 32, 64)), ('ffnn3 dgrad', LIN(B, 32, 64)), ('ffnn2 wgrad', LIN(B, 64, 128)), ('ffnn2 dgrad', LIN(B, 64, 128)),
         ('ffnn1 wgrad', LIN(B, 128, 256)), ('ffnn1 dgrad', LIN(B, 128, 256)), ('ffnn0 wgrad', LIN(B, 256, 562))]
_CNN = [('conv3 wgrad', CONV(25, 256, 512)), ('conv3 dgrad', CONV(25, 256, 512)), ('conv2 wgrad', CONV(58, 96, 256)),
        ('conv2 dgrad', CONV(58, 96, 256)), ('conv1 wgrad', CONV(124, 64, 96)), ('conv1 dgrad', CONV(124, 64, 96))]
# a forked step (the default: engine.cu backward_impl) issues the FFNN stack's backward on its side stream BEFORE docking_1's; the
# serial step (EMB_FORK=0, or the eager profile pass) issues it after.  ncu lists launches in issue order, so the table has to know
# which of the two it is looking at: the order under which no launch beats the burst peak is the right one (checked below).
BWD_ORDERS = {'forked': _HEAD + _FFNN + _DOCK1 + _CNN, 'serial': _HEAD + _DOCK1 + _FFNN + _CNN}

def main(path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peaks = {}
    if os.path.exists(os.path.join(root, 'MEASURED_PEAKS.json')):
        peaks = json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))
    tf_peak = peaks.get('bf16_tflops_sustained', 1382.0)
    bw_peak = peaks.get('hbm_gbs', 6547.0)
    recs = load(path)
    T = 'gpu__time_duration.sum'
    firsts = [i for i, d in enumerate(recs) if d['name'].startswith('wcache_fused_kernel')]
    assert len(firsts) >= 2, 'need at least one complete step (two weight-cache refreshes) in the capture'
    lo, hi = firsts[-2], firsts[-1]                                    # the last complete step of the capture
    step = recs[lo:hi]
    gemms = [d for d in step if d['name'] in GEMM]
    assert len(gemms) == len(FWD) + len(BWD_ORDERS['serial']), len(gemms)
    burst = peaks.get('bf16_tflops', 1607.2)
    fits = [k for k, bwd in BWD_ORDERS.items() if all(fl / d[T] / 1e6 <= 1.05 * burst for (_, fl), d in zip(FWD + bwd, gemms))]
    assert len(fits) == 1, ('cannot tell the issue order of this capture', fits)
    BWD = BWD_ORDERS[fits[0]]
    print(f'issue order of the captured step: {fits[0]}')
    print(f'peaks: {tf_peak:.0f} TFLOP/s bf16 (sustained, measured), {bw_peak:.0f} GB/s HBM (measured)\n')
    print(f"{'GEMM-class launch':34s}{'kernel':24s}{'us':>9s}{'GFLOP':>10s}{'TFLOP/s':>10s}{'frac':>7s}")
    tot_t = tot_f = 0.0
    for (label, fl), d in zip(FWD + BWD, gemms):
        t = d[T]
        tot_t += t; tot_f += fl
        print(f'{label:34s}{d["name"]:24s}{t:9.1f}{fl / 1e9:10.2f}{fl / t / 1e6:10.1f}{fl / t / 1e6 / tf_peak:7.2f}')
    print(f'{"GEMM class":58s}{tot_t:9.1f}{tot_f / 1e9:10.2f}{tot_f / tot_t / 1e6:10.1f}{tot_f / tot_t / 1e6 / tf_peak:7.2f}\n')
    print(f"{'other kernels (per class)':58s}{'us':>9s}{'DRAM MB':>10s}{'GB/s':>10s}{'frac':>7s}")
    agg = {}
    for d in step:
        if d['name'] in GEMM:
            continue
        a = agg.setdefault(d['name'], [0, 0.0, 0.0])
        a[0] += 1; a[1] += d[T]; a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    for name, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{(name + f" x{n}")[:57]:58s}{t:9.1f}{b / 1e6:10.1f}{b / t / 1e3:10.0f}{b / t / 1e3 / bw_peak:7.2f}')
    print(f'\nstep under ncu: {sum(d[T] for d in step):.1f} us in {len(step)} launches')


if __name__ == '__main__':
    main(sys.argv[1])
