"""Per-source-line executed warp instructions of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).
usage: python tools/ncu_instcount.py report.ncu-rep kernel_regex [top_n]"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per, thr, fname, hdr, nk = collections.Counter(), collections.Counter(), '', None, 0
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = {c: i for i, c in enumerate(r)}
        nk += 1
    elif hdr and r and r[0].isdigit() and len(r) > hdr['Instructions Executed']:
        try:
            v, t = float(r[hdr['Instructions Executed']]), float(r[hdr['Thread Instructions Executed']])
        except ValueError:
            continue
        key = (f'{fname}:{r[0]}', r[1].strip()[:100])
        per[key] += v
        thr[key] += t
tot = sum(per.values())
print(f'== {kern}: {tot:.0f} warp instructions over {nk} source tables')
for k, v in per.most_common(top):
    print(f'{100 * v / tot:5.1f}%  lanes {thr[k] / max(v, 1):4.1f}  {k[0]:>22s}  {k[1]}')
