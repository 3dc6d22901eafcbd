"""Rank CUDA source lines of an .ncu-rep by warp-stall samples:  python tools/ncu_lines.py rep [kernel-id] [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
hdr = rows[hi]; si = hdr.index('# Samples')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
inst_i = hdr.index('Instructions Executed')
cur = None; lines = []
for r in rows:
    if r and r[0] == 'File Name': cur = r[1].split('/')[-1]
    if len(r) == len(hdr) and r[0] not in ('', 'Line No'):
        try: n = float(r[si])
        except ValueError: continue
        st = sorted(((float(r[i]) if r[i] not in ('-', '') else 0, h[6:]) for i, h in stall), reverse=True)[:3]
        try: ie = float(r[inst_i])
        except ValueError: ie = 0
        lines.append((n, cur, r[0], r[1].strip(), st, ie))
tot = sum(l[0] for l in lines); toti = sum(l[5] for l in lines)
print(f"total samples {tot:.0f}, warp-instructions executed {toti:.0f}")
for n, fn, ln, src, st, ie in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{100*n/tot:5.1f}% inst={100*ie/toti:4.1f}% {fn}:{ln:>4s} {src[:78]:78s} {[(c, int(v)) for v, c in st if v > 0]}")
