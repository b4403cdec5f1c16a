"""Summarises an `ncu --metrics ... --csv` launch list (one row per kernel launch and metric) into one line per launch."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows:
    d.setdefault((r[ii], r[ki].split('(')[0].replace('void ', '').replace('hvo::', '')), {})[r[mi]] = float(r[vi].replace(',', ''))
d = collections.OrderedDict((k, v) for k, v in d.items() if k[1].startswith('k_'))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T = sum(v['gpu__time_duration.sum'] for v in d.values())
I = sum(v['smsp__inst_executed.sum'] for v in d.values())
print(f'{len(d)} launches, {T / 1e6:.2f} ms serialised, {I / 1e9:.2f} G warp instructions ({I / B / 1e6:.2f} M per frame at batch {B})')
print(f'{"kernel":22s} {"ms":>8s} {"share":>6s} {"Mwarp-inst/frame":>17s} {"thr/warp":>8s} {"issue%":>7s} {"warps%":>7s} {"DRAM R+W MB/frame":>18s} {"L2 sect/frame":>14s}')
for k, v in d.items():
    dr = (v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)) / B / 1e6
    print(f"{k[1]:22s} {v['gpu__time_duration.sum'] / 1e6:8.3f} {100 * v['gpu__time_duration.sum'] / T:5.1f}% {v['smsp__inst_executed.sum'] / B / 1e6:17.3f} "
          f"{v['smsp__thread_inst_executed.sum'] / max(v['smsp__inst_executed.sum'], 1):8.1f} {v['smsp__issue_active.avg.pct_of_peak_sustained_elapsed']:7.1f} "
          f"{v.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):7.1f} {dr:18.3f} {v['lts__t_sectors.sum'] / B / 1e3:13.1f}K")
