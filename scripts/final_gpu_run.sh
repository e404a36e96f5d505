#!/bin/bash
# One-GPU evidence run of a round: full GPU test suite, smoke, default bench (both arms), ncu launch list + full capture of the GEMMs.
#   gpurun --timeout 1500 -- scripts/final_gpu_run.sh <tag>        then: python scripts/make_profiles.py gpurun_out/<tag>_bench.json ...
tag=${1:-final}
o=gpurun_out
python -m pytest tests -m gpu -q > $o/${tag}_gpu_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
B="bench.py --steps 8 --warmup 8 --no-cpu-baseline --no-eval --no-sustained --no-ingest --also """
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 48 -c 16 -o $o/prof_r02_cfg2 -f python $B > $o/${tag}_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $o/launches_r02_cfg2.csv python $B > $o/${tag}_ncu_launch.log 2>&1
