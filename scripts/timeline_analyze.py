"""Summarise gpurun_out/timeline.csv (scripts/timeline.py): per-kernel totals of ONE graph-replayed step, concurrency histogram,
idle time and where in the step the time goes (analysis only - taken under a profiler, never a bench number)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
rows = [(r["name"], float(r["start_us"]), float(r["dur_us"])) for r in csv.DictReader(open(path))]
rows.sort(key=lambda r: r[1])
# keep the LAST step: from the last cast_f32_bf16 of the flat buffer ... simpler: split at the largest gap
gaps = sorted(((rows[i + 1][1] - (rows[i][1] + rows[i][2]), i) for i in range(len(rows) - 1)), reverse=True)
cut = gaps[0][1] + 1 if gaps and gaps[0][0] > 200 else 0
step = rows[cut:]
t0, t1 = step[0][1], max(s + d for _, s, d in step)
print("step span %.1f us, %d kernels" % (t1 - t0, len(step)))
ev = []
for n, s, d in step:
    ev.append((s, 1)); ev.append((s + d, -1))
ev.sort()
hist, cur, last = collections.Counter(), 0, t0
for t, k in ev:
    hist[cur] += t - last
    cur += k; last = t
print("time (us) with k kernels resident:", {k: round(v) for k, v in sorted(hist.items())})


def short(n):
    n = re.sub(r"^void ", "", n)
    return re.split(r"[<(]", n)[0]


tot = collections.defaultdict(lambda: [0, 0.0])
for n, s, d in step:
    tot[short(n)][0] += 1; tot[short(n)][1] += d
print("sum of kernel durations %.1f us" % sum(v[1] for v in tot.values()))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-40s %5d %9.1f us  avg %7.1f" % (k[:40], v[0], v[1], v[1] / v[0]))
# coarse phases: 1 ms bins, share of the bin covered by >= 1 kernel and the dominant kernel name
print("per-ms bins: busy fraction, mean concurrency, dominant kernel")
nb = int((t1 - t0) // 1000) + 1
for b in range(nb):
    lo, hi = t0 + 1000 * b, t0 + 1000 * (b + 1)
    inb = [(short(n), max(s, lo), min(s + d, hi)) for n, s, d in step if s < hi and s + d > lo]
    cov = sum(e - s for _, s, e in inb)
    dom = collections.Counter()
    for n, s, e in inb:
        dom[n] += e - s
    # busy: union length
    pts = sorted((s, e) for _, s, e in inb)
    busy, ce = 0.0, lo
    for s, e in pts:
        if e > ce:
            busy += e - max(s, ce); ce = e
    print("  %2d-%2d ms: busy %.2f conc %.2f n=%3d %s" % (b, b + 1, busy / 1000, cov / 1000, len(inb), ", ".join("%s %.0f%%" % (k[:22], v / 10) for k, v in dom.most_common(3))))
