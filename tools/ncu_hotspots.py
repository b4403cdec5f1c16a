"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).
usage: python tools/ncu_hotspots.py report.ncu-rep kernel_regex [top_n]"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per, why, fname = collections.Counter(), collections.defaultdict(collections.Counter), ''
hdr = None
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
        h = {c: i for i, c in enumerate(hdr)}
        ci = h['Warp Stall Sampling (All Samples)']
        stalls = [c for c in hdr if c.startswith('stall_') and '(' not in c]
    elif hdr and len(r) > ci and r[0].isdigit():   # a CUDA source line with the totals of its SASS rows
        try:
            v = float(r[ci])
        except ValueError:
            continue
        key = (f'{fname}:{r[0]}', r[1].strip())
        per[key] += v
        for s in stalls:
            try:
                why[key][s] += float(r[h[s]])
            except (ValueError, IndexError):
                pass
tot = sum(per.values())
print(f'== {kern}: {int(tot)} samples')
for k, v in per.most_common(top):
    w = ', '.join(f'{s[6:]} {100 * c / max(v, 1):.0f}%' for s, c in why[k].most_common(3))
    print(f'{100 * v / tot:5.1f}%  {k[0]:>22s}  {k[1][:100]:100s} [{w}]')
