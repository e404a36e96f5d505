"""Turn the ncu outputs of scripts/gpu_prof.sh into the small text summaries committed under profiles/.
   python scripts/summarize_ncu.py <tag> <round>      (reads gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep)"""
import csv, subprocess, sys, collections, io, os

tag, rnd = sys.argv[1], sys.argv[2]
out = []
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = [(r[ki], float(r[vi].replace(",", "")) / 1000.0, r[gi]) for r in data if len(r) > vi]
idx = [i for i, s in enumerate(seq) if "set_batch" in s[0]]
# steady-state graph-replayed steps: between consecutive set_batch launches that are followed by a full step
steps = [seq[a:b] for a, b in zip(idx[:-1], idx[1:]) if b - a >= 8]
# graph-replayed resident steps split the optimizer over two streams (two launches); the un-graphed profile_step and
# the host-buffer steps (load_batch) are skipped
graph_steps = [s for s in steps if sum("optimizer" in k[0] for k in s) == 2 and not any("load_batch" in k[0] for k in s)]
steps = graph_steps[len(graph_steps) // 2:] or steps[len(steps) // 2:len(steps) // 2 + 3] or steps[-1:]
out.append("# ncu --metrics gpu__time_duration.sum --clock-control none, `bench.py --steps 6 --warmup 3` (%s), 1x B200" % tag)
out.append("# per-launch device time of ONE steady-state training step (serialised, cold cache: compare SHARES)")
st = steps[0]
tot = sum(s[1] for s in st)
for name, us, grid in st:
    short = name.split("(")[0].replace("void ", "").replace("sb::", "")
    out.append("%-58s grid=%-14s %8.2f us  %5.1f%%" % (short[:58], grid, us, 100 * us / tot))
out.append("%-58s %22s %8.2f us" % ("TOTAL (sum of launches)", "", tot))
gemm = sum(s[1] for s in st if "gemm_tc" in s[0])
out.append("gemm_tc_kernel share of the step: %.1f%%" % (100 * gemm / tot))
open("profiles/ncu_r%s_%s_launches.txt" % (rnd, tag), "w").write("\n".join(out) + "\n")
print("\n".join(out))

rep = "gpurun_out/prof_%s.ncu-rep" % tag
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h = r[0]
    want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "smsp__cycles_active.avg"]
    have = [w for w in want if w in h]
    o2 = ["# ncu --set full --clock-control none -k regex:gemm_tc (two steady-state steps), %s, 1x B200" % tag, "# " + " | ".join(have),
          "# units: " + " | ".join(r[1][h.index(w)] for w in have)]
    for x in r[2:]:
        if len(x) < len(h):
            continue
        vals = [x[h.index(w)] for w in have]
        vals[0] = vals[0].split("(")[0].replace("void ", "").replace("sb::", "")[:44]
        o2.append(" | ".join(vals))
    open("profiles/ncu_r%s_%s_gemm_full.txt" % (rnd, tag), "w").write("\n".join(o2) + "\n")
    print("\n".join(o2[:12]))
