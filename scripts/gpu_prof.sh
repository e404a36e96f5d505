#!/bin/bash
# ncu evidence for one bench configuration.  Usage: bash scripts/gpu_prof.sh <tag> [bench args...]
tag=$1; shift
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/host_${tag}.txt 2>&1
import os
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("cpu.max n/a", e)
PY
# (1) every launch with its device time (serialised, cold cache: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also "" "$@" > gpurun_out/ncu_launch_${tag}.log 2>&1
# (2) full counters for the GEMM kernels of two steady-state steps
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 24 -c 16 -o gpurun_out/prof_${tag} -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also "" "$@" > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out | tail -5
