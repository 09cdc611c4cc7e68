"""Summarise an ncu --set full report (raw csv page) into one line per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name):
    i = col.get(name)
    return r[i] if i is not None else ''
print('kernel | grid | dur_us | dram_read_MB | dram_write_MB | dram_pct | sm_pct | tensor_pct | fp32_pipe_pct | regs')
for r in rows[2:]:
    def f(name, scale=1.0):
        v = g(r, name).replace(',', '')
        try: return float(v) * scale
        except ValueError: return float('nan')
    def unit(name): return units[col[name]] if name in col else ''
    dur = f('gpu__time_duration.sum'); u = unit('gpu__time_duration.sum')
    dur_us = dur * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1)
    def mb(name):
        v = f(name); u2 = unit(name)
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u2, 1)
    print('%s | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s' % (
        g(r, 'Kernel Name')[:48], g(r, 'launch__grid_size'), dur_us, mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum'),
        f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), f('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
        f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        f('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'), g(r, 'launch__registers_per_thread')))
