"""One line per kernel launch from an `ncu --page raw --csv` export (profiles/capture_r02.sh keeps the export, not the
report): the handful of metrics DESIGN.md quotes.
usage: python profiles/summarize_raw_csv.py gpurun_out/r02_ncu_step_raw.csv > profiles/r02_ncu_step.txt"""
import csv
import sys

COLS = [('Kernel Name', 'kernel', 38), ('Grid Size', 'grid', 15), ('gpu__time_duration.sum', 'us', 8),
        ('dram__bytes_read.sum', 'rd MB', 8), ('dram__bytes_write.sum', 'wr MB', 8),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2%', 6),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM%', 6),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tens%', 6),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%', 6),
        ('launch__registers_per_thread', 'regs', 5)]


def main(path):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if not l.startswith('==')))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) if c in hdr else None for c, _, _ in COLS]
    print(' '.join(('%-*s' if i < 2 else '%*s') % (w, n) for i, (_, n, w) in enumerate(COLS)))
    for r in rows[2:]:
        out = []
        for i, ((c, n, w), k) in enumerate(zip(COLS, idx)):
            v = r[k] if k is not None else ''
            if n == 'kernel':
                v = v.split('(')[0].replace('void ', '').replace('mpqe::<unnamed>::', '').replace('unnamed>::', '')[:w]
            elif n == 'us' and v:
                x = float(v.replace(',', ''))
                u = units[k]
                v = '%.1f' % (x * 1e3 if u == 'ms' else x / 1e3 if u == 'ns' else x * 1e6 if u == 's' else x)
            elif n in ('rd MB', 'wr MB') and v:
                x = float(v.replace(',', ''))
                u = units[k]
                v = '%.2f' % (x / 1e6 if u == 'byte' else x / 1e3 if u == 'Kbyte' else x * 1e3 if u == 'Gbyte' else x)
            elif v and n != 'grid':
                try:
                    v = '%.1f' % float(v.replace(',', ''))
                except ValueError:
                    pass
            out.append(('%-*s' if i < 2 else '%*s') % (w, v))
        print(' '.join(out))


if __name__ == '__main__':
    main(sys.argv[1])
