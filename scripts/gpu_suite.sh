#!/bin/bash
# Runs the GPU tests in separate processes (a trapped kernel poisons its CUDA context, so isolate), then the bench.
# Usage (on the GPU box, from the repo root):  bash scripts/gpu_suite.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/suite.log
  timeout $to python -m pytest -m gpu -q -x "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? $(tail -1 gpurun_out/$name.log)" | tee -a gpurun_out/suite.log
}
: > gpurun_out/suite.log
run gemm_identity 180 tests/test_gemm_tc.py -k identity
run gemm_shapes 300 tests/test_gemm_tc.py -k matches
run gemm_tiles 300 tests/test_gemm_tc.py -k every_tile
run trainer_fp32 300 tests/test_trainer_parity.py -k "fp32 or resident or epoch or async"
run trainer_bf16 300 tests/test_trainer_parity.py -k bf16
run scorer 300 tests/test_scorer_parity.py
run full_size 600 tests/test_full_size_properties.py
run host_mirrors 300 tests/test_host_mirrors.py
run text_ingest 300 tests/test_text_ingest.py
run multi_gpu 300 tests/test_multi_gpu.py
echo "=== smoke" | tee -a gpurun_out/suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/suite.log
echo "=== bench" | tee -a gpurun_out/suite.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" | tee -a gpurun_out/suite.log
tail -c 3000 gpurun_out/bench.log
