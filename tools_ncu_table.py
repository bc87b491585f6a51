"""Developer tool: per-kernel evidence table from `ncu --set full` captures exported on the GPU box with
`ncu -i X.ncu-rep --page raw --csv > X.raw.csv` (the reports themselves are too large to copy back).
  python tools_ncu_table.py gpurun_out/r2c_ncu_*.raw.csv > profiles/r2_ncu_kernel_table.md
  python tools_ncu_table.py --dominant-json profiles/r2_ncu_dominant.json gpurun_out/r2c_ncu_dominant.raw.csv
One row per captured launch: duration, DRAM bytes moved and achieved GB/s against the measured copy bandwidth
(MEASURED_PEAKS.json), tensor-pipe / FMA / XU (MUFU) / LSU utilisation, issue-slot utilisation, registers, grid."""
import argparse, csv, io, json, os, re, sys

ROOT = os.path.dirname(os.path.abspath(__file__))
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6554.2

COLS = [
    ("gpu__time_duration.sum", "us", lambda v, u: to_us(v, u)),
    ("dram__bytes_read.sum", "DRAM rd MB", lambda v, u: to_bytes(v, u) / 1e6),
    ("dram__bytes_write.sum", "DRAM wr MB", lambda v, u: to_bytes(v, u) / 1e6),
]
PCT = [
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
]


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return float("nan")


def to_us(v, u):
    x = num(v)
    return {"ns": x / 1e3, "us": x, "usecond": x, "ms": x * 1e3, "msecond": x * 1e3, "nsecond": x / 1e3, "s": x * 1e6, "second": x * 1e6}.get(u, x / 1e3)


def to_bytes(v, u):
    x = num(v)
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.replace("artalk::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")


def rows_of(path):
    txt = open(path).read()
    i = txt.find('"ID"')
    if i < 0:
        return
    r = list(csv.reader(io.StringIO(txt[i:])))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        if len(row) != len(hdr):
            continue
        yield dict(zip(hdr, row)), dict(zip(hdr, units))


ap = argparse.ArgumentParser()
ap.add_argument("files", nargs="+")
ap.add_argument("--dominant-json", default=None, help="write {launches: [{M,N,K,dram_bytes,us}]} for bench.py's roofline.traffic")
ap.add_argument("--dominant-shapes", default="149051x3072x1024,149051x1024x1024,149051x4096x1024,149051x1024x4096",
                help="M x N x K of the captured gemm_tc2 launches in capture order (layer 0: qkv, out-proj, ffn1, ffn2)")
a = ap.parse_args()

print("| kernel (capture) | grid x block | regs | us | DRAM rd MB | DRAM wr MB | GB/s | of %.0f | " % HBM + " | ".join(n for _, n in PCT) + " |")
print("|---|---|---:|---:|---:|---:|---:|---:|" + "---:|" * len(PCT))
dom = []
for f in a.files:
    tag = os.path.basename(f).replace(".raw.csv", "").replace("r2c_ncu_", "")
    for d, u in rows_of(f):
        us = to_us(d.get("gpu__time_duration.sum", "nan"), u.get("gpu__time_duration.sum", "ns"))
        rd = to_bytes(d.get("dram__bytes_read.sum", "nan"), u.get("dram__bytes_read.sum", "byte"))
        wr = to_bytes(d.get("dram__bytes_write.sum", "nan"), u.get("dram__bytes_write.sum", "byte"))
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else float("nan")
        cells = ["%.1f" % num(d.get(k, "nan")) for k, _ in PCT]
        print("| `%s` (%s) | %s x %s | %s | %.1f | %.2f | %.2f | %.0f | %.1f%% | " % (
            short(d.get("Kernel Name", "?"))[:70], tag, d.get("Grid Size", "?"), d.get("Block Size", "?"),
            d.get("launch__registers_per_thread", "?"), us, rd / 1e6, wr / 1e6, gbs, 100 * gbs / HBM) + " | ".join(cells) + " |")
        dom.append({"kernel": short(d.get("Kernel Name", "?")), "us": us, "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                    "tensor_pct": num(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "nan"))})
if a.dominant_json:
    shapes = [tuple(int(x) for x in s.split("x")) for s in a.dominant_shapes.split(",")]
    out = {"source": "ncu --set full --clock-control none, `python bench.py --profile-step --steps 1 --warmup 0`, gemm_tc2_kernel launches "
                     "8-11 of the step (wav2vec layer 0), " + os.path.basename(a.files[0]), "launches": []}
    for e, (M, N, K) in zip(dom, shapes):
        out["launches"].append(dict(e, M=M, N=N, K=K))
    json.dump(out, open(a.dominant_json, "w"), indent=1)
