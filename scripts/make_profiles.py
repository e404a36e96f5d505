#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (committed evidence), round 2.

    python scripts/make_profiles.py <bench_json> [<bench_json_n2> ...]

  profiles/ncu_r02_cfg2_gemm_full.txt   per-kernel table of the `ncu --set full` capture (gpurun_out/prof_r02_cfg2.ncu-rep)
  profiles/ncu_r02_cfg2_launches.txt    launch list with device times (gpurun_out/launches_r02_cfg2.csv)
  profiles/ncu_r02_traffic.json         dram bytes per launch, keyed "<config>/<kernel role>" (bench.py roofline.traffic)
  profiles/parity_r02.json              observed errors of the parity tests on the benchmarked paths
  profiles/bench_r02_*.json             the bench lines the results table quotes
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
ROLES = ["fwd0", "fwd1", "fwd2+out", "dW2", "dA2", "dW1", "dA1", "dW0"]      # launch order of the GEMMs of one cfg2 step

rep = "gpurun_out/prof_r02_cfg2.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h = r[0]
    want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "smsp__cycles_active.avg", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    have = [w for w in want if w in h]
    out = ["# ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel (two steady-state cfg2 steps of bench.py), 1x B200",
           "# (cold-cache, serialised replays: compare shares and counters, not absolute times; in-step times are bench.py's roofline.kernels)",
           "# role | " + " | ".join(have), "# units: " + " | ".join(r[1][h.index(w)] for w in have)]
    traffic = {}
    k = 0
    for x in r[2:]:
        if len(x) < len(h):
            continue
        vals = [x[h.index(w)] for w in have]
        vals[0] = vals[0].split("(")[0].replace("void ", "").replace("sb::", "")[:52]
        role = ROLES[k % len(ROLES)]
        out.append(role + " | " + " | ".join(vals))
        if k < len(ROLES):
            unit = r[1][h.index("dram__bytes_read.sum")]
            mul = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}.get(unit, 1e6)
            traffic["cfg2/" + role] = (float(x[h.index("dram__bytes_read.sum")]) + float(x[h.index("dram__bytes_write.sum")])) * mul
        k += 1
    open("profiles/ncu_r02_cfg2_gemm_full.txt", "w").write("\n".join(out) + "\n")
    json.dump(traffic, open("profiles/ncu_r02_traffic.json", "w"), indent=1, sort_keys=True)
    print("\n".join(out[:14]))

lc = "gpurun_out/launches_r02_cfg2.csv"
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hi = [i for i, x in enumerate(rows) if x and x[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = [(x[ki], float(x[vi].replace(",", "")) / 1000.0, x[gi]) for x in data if len(x) > vi]
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none (bench.py, cfg2), 1x B200: per-launch device time,",
           "# serialised and cold-cache - compare SHARES; one steady-state step = the launches between two set_batch_kernel launches"]
    idx = [i for i, s in enumerate(seq) if "set_batch" in s[0]]
    steps = [seq[a:b] for a, b in zip(idx[:-1], idx[1:]) if b - a >= 8]
    st = steps[len(steps) // 2] if steps else seq[:12]
    tot = sum(s[1] for s in st)
    for name, us, grid in st:
        short = name.split("(")[0].replace("void ", "").replace("sb::", "")
        out.append("%-62s grid=%-14s %8.2f us  %5.1f%%" % (short[:62], grid, us, 100 * us / tot))
    out.append("%-62s %22s %8.2f us" % ("TOTAL (sum of launches)", "", tot))
    out.append("gemm_tc_kernel share of the step: %.1f%%" % (100 * sum(s[1] for s in st if "gemm_tc" in s[0]) / tot))
    open("profiles/ncu_r02_cfg2_launches.txt", "w").write("\n".join(out) + "\n")

if os.path.exists("gpurun_out/parity_benchmarked_paths.json"):
    shutil.copy("gpurun_out/parity_benchmarked_paths.json", "profiles/parity_r02.json")
for path in sys.argv[1:]:
    if os.path.exists(path) and os.path.getsize(path) > 0:
        d = next(json.loads(l) for l in open(path) if l.lstrip().startswith("{"))      # (torchrun may print around the JSON line)
        name = "profiles/bench_r02_n%d_%s.json" % (d.get("n_gpus", 1), d["config"]["workload"].split(":")[0])
        json.dump(d, open(name, "w"), indent=1)
        print("wrote", name)
