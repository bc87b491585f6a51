"""Developer tool: one markdown row of the headline ncu metrics per kernel in a ``--set full`` report.
  python tools_ncu_summary.py gpurun_out/prof_x.ncu-rep [...]"""
import csv, subprocess, sys, io
WANT = [("gpu__time_duration.sum", "duration"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("sm__cycles_active.avg", "SM active cycles"), ("smsp__inst_executed.sum", "warp instr")]
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").replace("artalk::<unnamed>::", "")[:60]
        cells = ["%s %s" % (d.get(k, "?"), u.get(k, "")) for k, _ in WANT]
        print("| `%s` (%s) | " % (name, path.split("/")[-1]) + " | ".join(cells) + " |")
print("columns: " + " | ".join(n for _, n in WANT))
