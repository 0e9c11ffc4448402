#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX OBJECT.o [MANGLED_SUBSTR]
Per-source-line view of an ncu --set full capture without the GUI: SASS offsets of the kernel
(ncu --page source --csv) are joined with the line table of the same build (nvdisasm -g) and the
executed warp instructions, active lanes and stall samples are summed per file:line."""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kre, obj = sys.argv[1:4]
sub = sys.argv[4] if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
kname = rows[0][1]
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # (a second matching launch)
    if len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
        data.append(r)
ix = {n: i for i, n in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
# pick the function: template arguments of the ncu name -> mangled ILb..E fragment, or explicit substring
if sub is None:
    targs = re.search(r"<(.*?)>\(", kname)
    base_name = re.search(r"(\w+)<", kname).group(1) if targs else re.search(r"(\w+)\(", kname).group(1)
    frag = base_name
    if targs:
        frag += "I" + "".join(("Li%dE" if "(int)" in a else "Lb%dE") % int(a.split(")")[-1]) for a in targs.group(1).split(",")) + "E"
    sub = frag
line_of, cur, inside = {}, None, False
for l in sass:
    if l.startswith(".text.") and l.rstrip().endswith(":"):
        inside = sub in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur


def f(r, n):
    try:
        return float(r[ix[n]])
    except (KeyError, ValueError):
        return 0.0


agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
miss = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off)
    if key is None:
        miss += 1
        key = ("?", 0)
    a = agg[key]
    a[0] += f(r, "Instructions Executed"); a[1] += f(r, "Thread Instructions Executed")
    a[2] += f(r, "# Samples"); a[3] += f(r, "stall_long_sb")
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"{kname}: {len(data)} SASS instructions ({miss} unmapped), {ti:.3g} warp instructions, "
      f"{sum(a[1] for a in agg.values()) / ti:.1f} lanes on average")
src = {}
print(f"{'file:line':28s} {'%inst':>6s} {'lanes':>6s} {'%smpl':>6s} {'%lsb':>6s}  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2 if os.environ.get("SORT") == "samples" else 0])[:int(os.environ.get("TOP", "45"))]:
    fn, ln = key
    if fn not in src:
        for d in ("flexpart_b200/csrc", "."):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, fn)
            if os.path.exists(p):
                src[fn] = open(p).read().splitlines()
                break
        else:
            src[fn] = []
    text = src[fn][ln - 1].strip()[:90] if 0 < ln <= len(src[fn]) else ""
    print(f"{fn + ':' + str(ln):28s} {100 * a[0] / ti:6.2f} {a[1] / max(a[0], 1):6.1f} {100 * a[2] / ts:6.2f} {100 * a[3] / ts:6.2f}  {text}")
