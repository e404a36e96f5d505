#!/bin/bash
# One-GPU evidence run of a round: default bench, ncu launch list + full capture of the GEMMs, smoke, full GPU test suite (last:
# the GPU budget may end the call early).
#   gpurun --timeout 900 -- scripts/final_gpu_run.sh <tag>        then: python scripts/make_profiles.py gpurun_out/<tag>_bench.json ...
tag=${1:-final}
o=gpurun_out
timeout 300 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
B="bench.py --steps 8 --warmup 8 --no-cpu-baseline --no-eval --no-sustained --no-ingest --e2e-steps 3 --also none"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 48 -c 16 -o $o/prof_r02_cfg2 -f python $B > $o/${tag}_ncu_full.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $o/launches_r02_cfg2.csv python $B > $o/${tag}_ncu_launch.log 2>&1
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1
timeout 420 python -m pytest tests -m gpu -q -x > $o/${tag}_gpu_tests.log 2>&1
