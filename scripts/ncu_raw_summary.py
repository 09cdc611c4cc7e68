"""One line per captured launch from an `ncu --page raw --csv` export (the .ncu-rep itself is not kept: gpurun_out is capped)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], newline='')))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name, scale_map=None):
    i = col.get(name)
    if i is None: return float('nan')
    try: v = float(r[i].replace(',', ''))
    except ValueError: return float('nan')
    if scale_map: v *= scale_map.get(units[i], 1.0)
    return v
T = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
B = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
print('kernel | grid | block | dur_us | dram_read_MB | dram_write_MB | dram_pct | sm_pct | tensor_active_pct | issue_active_pct | regs | sm_mhz')
for r in rows[2:]:
    name = r[col['Kernel Name']].replace('(anonymous namespace)::', '').replace('db::', '').split('(')[0][-44:]
    print('%s | %s | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %.0f' % (
        name, r[col.get('launch__grid_size', 0)], r[col.get('launch__block_size', 0)], val(r, 'gpu__time_duration.sum', T),
        val(r, 'dram__bytes_read.sum', B), val(r, 'dram__bytes_write.sum', B),
        val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
        val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        val(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active'), r[col.get('launch__registers_per_thread', 0)],
        val(r, 'sm__cycles_elapsed.avg.per_second', {'cycle/nsecond': 1e3, 'cycle/usecond': 1.0, 'Ghz': 1e3, 'Mhz': 1.0, 'hz': 1e-6})))
