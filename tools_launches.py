"""Developer tool: summarise an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv``) of ``bench.py`` into a
markdown table of per-kernel launches / total time / share for ONE step.
  python tools_launches.py gpurun_out/launches.csv [--steps-in-file 5] > profiles/rXX_launches_summary.md
The bench runs (warmup + steps) + e2e (1 + steps) + 1 instrumented step; the last step in the file (the instrumented one,
graphs off) and the ones before it launch the same kernels, so the script takes the last complete step: launches between the
last two ``audio_stats_kernel`` launches ... end of file."""
import csv, sys, re, collections, argparse

ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("--which", type=int, default=-2, help="index of the step to summarise (python index into the list of steps)")
ap.add_argument("--detail", action="store_true")
ap.add_argument("--all", action="store_true", help="the file holds exactly one step (bench.py --profile-step --steps 1 --warmup 0): summarise all of it")
a = ap.parse_args()
rows = []
for r in csv.reader(l for l in open(a.csv) if l.startswith('"')):
    if r[0] == "ID":
        hdr = r; continue
    rows.append(r)
ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
names = [r[ik] for r in rows]
starts = [i for i, n in enumerate(names) if "audio_stats_kernel" in n]
bounds = starts + [len(rows)]
steps = [(bounds[i], bounds[i + 1]) for i in range(len(starts))]
s0, s1 = (0, len(rows)) if a.all else steps[a.which]
def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.replace("artalk::", "").replace("(anonymous namespace)::", "")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[s0:s1]:
    k = short(r[ik])
    if a.detail:
        k = k + " grid" + r[ig]
    us = float(r[iv].replace(",", "")) / 1e3
    e = agg.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += us; tot += us
print("steps found: %d; summarising step %d: %d launches, %.2f ms summed\n" % (len(steps), a.which, s1 - s0, tot / 1e3))
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.2f | %.1f%% | %.1f |" % (k, c, us / 1e3, 100 * us / tot, us / c))
