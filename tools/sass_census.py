"""Per-kernel SASS opcode census of the shipped library: tcgen05 (UTC*MMA), TMEM loads/stores (LDTM/STTM), TMA
(UTMALDG/UTMASTG, UBLKCP), warp-level MMA (HMMA), MUFU.   python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "instancediff_b200", "libidiff_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", root], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "SYNCS", "MUFU", "LDGSTS"]
name, cnt, rows = None, None, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, cnt))
        name, cnt = m.group(1), collections.Counter()
        continue
    if name:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            cnt["total"] += 1
            for p in pats:
                if m.group(1).startswith(p):
                    cnt[p] += 1
if name:
    rows.append((name, cnt))
tot = collections.Counter()
print(f"{'kernel':78s} " + " ".join(f"{p:>8s}" for p in pats + ["total"]))
for n, c in rows:
    dem = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("idiff::", "")
    dem = re.sub(r"\(.*", "", dem)
    print(f"{dem[:78]:78s} " + " ".join(f"{c[p]:8d}" for p in pats + ["total"]))
    tot.update(c)
print(f"{'ALL KERNELS (' + str(len(rows)) + ')':78s} " + " ".join(f"{tot[p]:8d}" for p in pats + ["total"]))
