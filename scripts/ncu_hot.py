"""Stall samples of one kernel of an .ncu-rep grouped by how often the SASS line executed
(a proxy for the warp role) and the top lines.  usage: ncu_hot.py report.ncu-rep kernel"""
import collections, csv, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", k], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((int(r[isamp]), r[ia].strip(), int(r[iex]), i))
    except Exception:
        pass
seen, uniq = set(), []
for d in data:                      # the listing repeats per launch: keep the first
    if d[1:3] + (d[0],) in seen and False:
        continue
    uniq.append(d)
n_launch = max(1, sum(1 for r in rows if r and r[0] == "Kernel Name"))
data = uniq[:len(uniq) // n_launch]
tot = sum(d[0] for d in data)
by, n = collections.Counter(), collections.Counter()
for s, src, ex, i in data:
    by[ex] += s; n[ex] += 1
print("total samples", tot, "launches in report", n_launch)
for ex, s in sorted(by.items(), key=lambda x: -x[1])[:8]:
    print("ex=%9d lines=%4d samples=%6d %5.1f%%" % (ex, n[ex], s, 100 * s / tot))
for s, src, ex, i in sorted(data, reverse=True)[:14]:
    print("%6d %5.1f%%  ex=%8d  #%d  %s" % (s, 100 * s / tot, ex, i, src[:100]))
