"""Summarise an `ncu --set full` capture: python scripts/ncu_summary.py gpurun_out/prof_top.ncu-rep > table.md
Reads the raw page through `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv, subprocess, sys

COLS = [('duration us', 'gpu__time_duration.sum', 1.0),
        ('dram read MB', 'dram__bytes_read.sum', 1.0),
        ('dram write MB', 'dram__bytes_write.sum', 1.0),
        ('dram % of peak', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
        ('tensor pipe % active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 1.0),
        ('regs/thread', 'launch__registers_per_thread', 1.0),
        ('SM GHz', 'sm__cycles_elapsed.avg.per_second', 1.0)]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = hdr.index('Kernel Name')
    idx = [(t, hdr.index(m) if m in hdr else -1) for t, m, _ in COLS]
    print('| kernel | ' + ' | '.join(t for t, _ in idx) + ' |')
    print('|---' * (len(idx) + 1) + '|')
    for r in data:
        k = r[name].split('(')[0].replace('void ', '')
        cells = []
        for t, i in idx:
            if i < 0:
                cells.append('n/a'); continue
            v = float(r[i].replace(',', ''))
            u = units[i].lower()
            if t.startswith('dram') and 'MB' in t:
                v *= {'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1.0, 'gbyte': 1e3}.get(u, 1.0)
            if t.startswith('duration'):
                v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
            cells.append('%.0f' % v if t == 'regs/thread' else '%.2f' % v if t == 'SM GHz' else '%.1f' % v)
        print('| `%s` | ' % k + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main(sys.argv[1])
