"""Join an ncu SASS-level sample dump with nvdisasm line info: hot source lines of one kernel.

  ncu -i X.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all file.o ; nvdisasm -g file.cubin > dis.txt
  python profiles/tools/sass_hotlines.py sass.csv dis.txt '<section name substring>' [top]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis_txt, section = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file, line) from nvdisasm -g
addr_line = {}
cur, active = None, False
for ln in open(dis_txt, errors="replace"):
    if ln.startswith("//--------------------- .text."):
        active = section in ln
        cur = None
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
first = None
hdr = None
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        if first is None:
            first = a
        a -= first          # ncu prints absolute addresses; nvdisasm offsets from the start of the function
        s = int(r[hdr.index("# Samples")])
        ie = int(r[hdr.index("Instructions Executed")])
    except Exception:
        continue
    base = min(addr_line) if addr_line else 0
    key = addr_line.get(a) or addr_line.get(a - (a - base) % 16) or ("?", 0)
    agg[key][0] += s
    agg[key][1] += ie
    for st in [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]:
        try:
            agg[key][2][st] += int(r[hdr.index(st)])
        except Exception:
            pass
    tot += s
print("total samples", tot)
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, line = k
    if f not in src:
        try:
            import glob
            p = [x for x in glob.glob("yoloseries_b200/csrc/*") if x.endswith(f)]
            src[f] = open(p[0]).read().split("\n") if p else []
        except Exception:
            src[f] = []
    text = src[f][line - 1].strip()[:90] if src[f] and 0 < line <= len(src[f]) else ""
    stalls = sorted(v[2].items(), key=lambda kv: -kv[1])[:2]
    print(f"{100 * v[0] / max(tot, 1):5.1f}% {f}:{line:4d} inst={v[1]:8d} {','.join(f'{a[6:]}={b}' for a, b in stalls):32s} {text}")
