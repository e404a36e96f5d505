#!/usr/bin/env python
"""profiles/bench_r02_*.json (+ the variant runs still in gpurun_out/) -> profiles/results_r02.md

    python scripts/make_profiles.py gpurun_out/<n1>.json gpurun_out/<n2>.json gpurun_out/<n8>.json && python scripts/make_results.py
"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def load(path):
    if not os.path.exists(path):
        return None
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    try:
        return json.load(open(path))
    except ValueError:
        return None


def krow(k):
    return "| %s | %.2f | %.2f | %.2f | %s |" % (k["kernel"], k["begin_us"], k["end_us"], k["us"], ("%.0f" % k["tflops"]) if k.get("tflops") else "")


out = ["# Round 2 results (B200, this pool; every number from a `gpurun` call of this round, files named per line)", ""]
b = {n: load("profiles/bench_r02_n%d_cfg2.json" % n) for n in (1, 2, 4, 8)}
n1 = b[1]
if n1:
    r = n1["roofline"]
    out += ["## 1 GPU, cfg2 (2000 cols × 8192 rows, MLP [1024, 512, 256], momentum, bf16) — `profiles/bench_r02_n1_cfg2.json`", "",
            "| | |", "|---|---|",
            "| `value` (resident, burst) | **%.2f M rows/s**, %.4f ms/step |" % (n1["value"] / 1e6, n1["ms_per_step"]),
            "| whole step vs measured cuBLAS bf16 peak (%.0f TF/s) | %.3f |" % (r["peak"], r.get("step_fraction_of_peak", 0))]
    if n1.get("sustained"):
        s = n1["sustained"]
        out.append("| sustained (%.1f s, clocks sampled) | %.2f M rows/s, %.4f ms/step, %.3f of the sustained peak (%.0f TF/s); clocks %s |" % (
            s["seconds"], s["value"] / 1e6, s["ms_per_step"], s["step_fraction_of_sustained_peak"], s["peak"], json.dumps(s.get("clocks"))))
    out.append("| longest GEMM (`roofline`) | %s: %.1f µs, %.0f TF/s = %.3f of peak; DRAM traffic per launch %s |" % (
        r.get("kernel", ""), r.get("kernel_us", 0), r["achieved"], r["frac"], r.get("traffic")))
    if r.get("all_gemms"):
        out.append("| all GEMMs of the step | %.1f GFLOP in %.1f µs of kernel time = %.0f TF/s (%.3f) |" % (
            r["all_gemms"]["flops"] / 1e9, r["all_gemms"]["sum_kernel_us"], r["all_gemms"]["tflops"], r["all_gemms"]["frac"]))
    e = n1.get("e2e")
    if e:
        out.append("| `e2e` (host buffers, H2D + loss D2H inside) | %.2f M rows/s (%.1f MB H2D per step: PCIe bound) |" % (e["value"] / 1e6, e["h2d_bytes_per_step"] / 1e6))
    c = n1.get("cpu_baseline")
    if c:
        out.append("| `cpu_baseline` (%s, %s cores) | %.0f rows/s (%s) |" % (c.get("kind"), c.get("cores"), c["value"], c.get("sample")))
    ev = n1.get("eval")
    if ev:
        out.append("| eval (BASELINE config 5) | %s |" % json.dumps({k: ev[k] for k in ev if k in ("value", "unit", "rows", "e2e", "parity_modes", "roofline")})[:600])
    ig = n1.get("ingest")
    if ig:
        out.append("| ingest (text → fp32 columns on the device) | %s |" % json.dumps(ig)[:500])
    a = n1.get("also")
    if a:
        out.append("| cfg1 (1000 × 4096, [512, 256, 128], Adam) | %.2f M rows/s, %.4f ms/step (step %.3f of peak) |" % (
            a["value"] / 1e6, a["ms_per_step"], a["roofline"].get("step_fraction_of_peak", 0)))
    out += ["", "In-graph kernel spans of one cfg2 step (µs; `%globaltimer` stamps inside the captured graph, no profiler):", "",
            "| kernel | begin | end | span | TF/s |", "|---|---|---|---|---|"] + [krow(k) for k in r.get("kernels", [])] + [""]
    if a and a["roofline"].get("kernels"):
        out += ["cfg1 step:", "", "| kernel | begin | end | span | TF/s |", "|---|---|---|---|---|"] + [krow(k) for k in a["roofline"]["kernels"]] + [""]

out += ["## Scaling (weak, cfg2, one NVSwitch node) — `profiles/bench_r02_n{1,2,4,8}_cfg2.json`", "",
        "| N | M rows/s | µs/step | efficiency vs N = 1 |", "|---|---|---|---|"]
for n in (1, 2, 4, 8):
    if b[n] and n1:
        out.append("| %d | %.2f | %.1f | %.2f |" % (n, b[n]["value"] / 1e6, b[n]["ms_per_step"] * 1e3, b[n]["value"] / n / n1["value"]))
out.append("")
for n in (2, 8):
    if b[n]:
        out += ["N = %d step (rank 0; xchg rows: A = every layer but hidden layer 0, B0 / B1 = row chunks of layer 0):" % n, "",
                "| kernel | begin | end | span | TF/s |", "|---|---|---|---|---|"] + [krow(k) for k in b[n]["roofline"].get("kernels", [])] + [""]

# variants measured while the exchange was built (gpurun_out/ is scratch: the numbers are copied here)
rows = []
for tag, what in (("r2f_bench_n2", "first version: segment B PDL-chained behind dW_0, pushed all-gather + sys fence"),
                  ("r2h_bench_n2", "chunked dW_0, dW_1 behind it, slot A deferred (executor started it late), push"),
                  ("r2i_bench_n2", "pull all-gather, dW_1 beside dW_0"), ("r2j_bench_n2_nodefer", "pull, batched loads, one fence per block, no deferral"),
                  ("r2l_5", "+ gpu-scope fence, no deferral"), ("r2n_3", "slot A on the main stream as dW_1's PDL dependent (flag-and-pull)"),
                  ("r2o_2", "dW_1 FIRST, flag-and-pull everywhere"), ("r2o_1", "dW_1 first, LL for the last chunk only"),
                  ("r2o_3", "dW_1 first, LL everywhere (default)"), ("r2p_2", "no chunking, single-GPU GEMM order, LL"),
                  ("r2q_1", "N = 8: default"), ("r2r_1", "N = 8: flag-and-pull everywhere"), ("r2r_2", "N = 8: dW_1 last + deferred slot A, LL"),
                  ("r2r_3", "N = 8: dW_1 last + deferred slot A, flag-and-pull")):
    d = load("gpurun_out/%s.json" % tag)
    if d:
        xs = ["%s %.0f" % (k["kernel"], k["us"]) for k in d["roofline"].get("kernels", []) if k["kernel"].startswith("xchg")]
        rows.append("| %s | %d | %.1f | %.2f | %s |" % (what, d["n_gpus"], d["ms_per_step"] * 1e3, d["value"] / 1e6, ", ".join(xs)))
if rows:
    out += ["## The exchange, step by step (cfg2; µs/step, M rows/s, spans of the exchange launches in µs)", "",
            "| variant | N | µs/step | M rows/s | exchange launches |", "|---|---|---|---|---|"] + rows + [""]

p = load("profiles/parity_r02.json")
if p:
    out += ["## Parity observed on the benchmarked paths (`tests/test_benchmarked_paths.py`, `profiles/parity_r02.json`)", "", "```", json.dumps(p, indent=1)[:3500], "```", ""]
out += ["## Targets of `north_star` against what was measured", "",
        "| target | measured | where |", "|---|---|---|",
        "| hidden-layer GEMMs >= 70 % of the bf16 tensor peak | layer-0 forward 1.19-1.22 PFLOP/s = 0.70-0.72 of the measured cuBLAS burst peak (1690 TF/s); dW_0 1.13 PFLOP/s = 0.67; "
        "the small GEMMs are epilogue / latency bound (dA_1 0.28, fwd_1 0.50, layers of 256-512 columns 0.12-0.18); all GEMMs of a step together 0.47 | spans above; "
        "`profiles/ncu_r02_cfg2_gemm_full.txt` (tensor-pipe active: fwd_0 64 %, dW_0 64 %, dA_1 26 %) |",
        "| >= 0.9 scaling efficiency at 8 GPUs on the 2000 x 8192 batch | 0.73 at N = 8, 0.81 at N = 2 (round 1: 0.69 at N = 8 on cfg2) | table above; DESIGN section 7 says what the rest is |",
        "| loss / gradients within 1e-4 (fp32), scores within 1e-5 | fp32 and fp32_tc modes: loss curves <= 3.3e-5 (cfg1) / 9e-7 (cfg2) over 20-30 steps, single-step gradients <= 5e-6; "
        "bf16 (the benchmarked mode) vs the bf16-emulating oracle 1.9e-4 / 1.4e-5 | parity block below, `tests/test_benchmarked_paths.py`, `tests/test_trainer_parity.py`, `tests/test_scorer_parity.py` |",
        "| reference CPU worker next to it | 92 k rows/s on 16 host cores (torch-CPU port of the worker loop; TF-1.x cannot be installed here) - the reference arm `bench.py --impl reference` measures the same loop | `cpu_baseline` |",
        "",
        "What did not move the dA epilogue (kept, documented in DESIGN section 6): shared-memory column sums, TMA-staged tiles both ways, "
        "A_{l-1} prefetched before the accumulator wait - 19.5-21 us before and after; sixteen epilogue warps: 19.7 -> 17.8 us.",
        "The ncu capture (`profiles/ncu_r02_cfg2_*.txt`, `profiles/ncu_r02_traffic.json`) is of the build one commit before dW_1 moved in front of dW_0 "
        "on one GPU: same kernels, launch order dW_2, dA_2, dW_1, dA_1, dW_0.", ""]
open("profiles/results_r02.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
