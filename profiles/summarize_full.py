"""Table of the per-launch metrics of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`:
python profiles/summarize_full.py raw.csv > profiles/rNN_ncu_full_step.txt"""
import csv
import re
import sys

COLS = [('time', 'gpu__time_duration.sum'), ('sm clock', 'sm__cycles_elapsed.avg.per_second'),
        ('dram read', 'dram__bytes_read.sum'), ('dram write', 'dram__bytes_write.sum'),
        ('tensor pipe active %', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('UMMA bf16 % of peak', 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed'),
        ('IPC', 'sm__inst_executed.avg.per_cycle_elapsed'), ('L2->SM read', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second'),
        ('L2 hit %', 'lts__t_sector_hit_rate.pct'), ('regs', 'launch__registers_per_thread'),
        ('dyn smem', 'launch__shared_mem_per_block_dynamic')]


def main(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    ki, gi = idx['Kernel Name'], idx['Grid Size']
    present = [(t, m) for t, m in COLS if m in idx]
    scale = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

    def val(r, m):
        return float(r[idx[m]].replace(',', '')) * scale[U[idx[m]]]
    print(f"{'kernel':48s}{'grid':>14s}{'dram TB/s':>12s}" + ''.join(f'{t:>22s}' for t, _ in present))
    for r in rows[2:]:
        name = re.sub(r'^void ', '', r[ki]).replace('emb::', '').split('(')[0]
        cells = []
        for _, m in present:
            v, u = r[idx[m]], U[idx[m]]
            try:
                v = f'{float(v.replace(",", "")):.4g}'
            except ValueError:
                pass
            cells.append(f'{v} {u}'[:21])
        tbs = (val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')) / val(r, 'gpu__time_duration.sum') / 1e12
        print(f'{name[:47]:48s}{r[gi].replace(" ", ""):>14s}{tbs:12.2f}' + ''.join(f'{c:>22s}' for c in cells))


if __name__ == '__main__':
    main(sys.argv[1])
