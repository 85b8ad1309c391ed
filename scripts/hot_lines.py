"""Per-source-line warp-instruction counts of one kernel launch from an .ncu-rep captured with
--import-source on (the `source` page):  python scripts/hot_lines.py REPORT KERNEL_REGEX [top]"""
import collections, csv, re, subprocess, sys

rep, pat = sys.argv[1], re.compile(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
names = [r[rows[0].index("Kernel Name")] for r in rows[2:]]
idx = next(i for i, n in enumerate(names) if pat.search(n))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
agg, tot, stot = collections.OrderedDict(), 0, 0
fname, h, ci, si = "", None, 0, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        ci, si = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or not fname or len(r) <= ci or not r[0].isdigit():
        continue
    try:
        v = int(r[ci])
    except ValueError:
        continue
    smp = int(r[si]) if r[si].isdigit() else 0
    a = agg.setdefault((fname, int(r[0])), [0, 0, r[1]])
    a[0] += v
    a[1] += smp
    tot += v
    stot += smp
print(f"kernel: {names[idx][:100]}\ntotal warp instructions: {tot}   stall samples: {stot}")
for (fn, ln), (v, smp, text) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{fn}:{ln:<5d} {v:10d} {100 * v / max(tot, 1):5.1f}%  samples {100 * smp / max(stot, 1):5.1f}%  {text.strip()[:100]}")
