#!/bin/bash
# usage: scripts/bench_variants.sh <tag> <n_gpus> "<ENV=1 ENV2=1>" ["<more env>" ...]
# One N-GPU bench.py run (cfg2 only, resident + traced legs) per environment set; results in gpurun_out/<tag>_<i>.json
tag=$1; shift; n=$1; shift
export SB_XCHG_TIMEOUT_S=${SB_XCHG_TIMEOUT_S:-20}
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "variant $i: $envs" > "gpurun_out/${tag}_${i}.err"
  if [ "$n" -gt 1 ]; then
    env $envs timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600+i)) \
      bench.py --gpus "$n" --no-cpu-baseline --no-ingest --no-eval --no-sustained --e2e-steps 5 --also "" > "gpurun_out/${tag}_${i}.json" 2>> "gpurun_out/${tag}_${i}.err"
  else
    env $envs timeout 240 python bench.py --no-cpu-baseline --no-ingest --no-eval --no-sustained --e2e-steps 5 --also "" > "gpurun_out/${tag}_${i}.json" 2>> "gpurun_out/${tag}_${i}.err"
  fi
done
