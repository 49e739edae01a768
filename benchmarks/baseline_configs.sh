#!/bin/bash
# The BASELINE.json configurations beyond the headline, end to end through bench.py on the GPUs of this box (default 8).
# Usage: bash benchmarks/baseline_configs.sh [N] [out.jsonl]   -> one JSON line per configuration
N=${1:-8}
OUT=${2:-gpurun_out/baseline_configs.jsonl}
mkdir -p "$(dirname $OUT)"
: > $OUT
run() {
  if [ "$N" -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $N --steps 20 --warmup 5 "$@"
  else
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 "$@"
  fi 2>&1 | grep -a '^{"metric' | tee -a $OUT | cut -c1-260
}
run --experiment cnnet --aggregator krum --nb-workers 8 --nb-decl-byz-workers 2 --no-baseline                      # configuration 2
run --aggregator bulyan --nb-workers 16 --nb-decl-byz-workers 2                                                    # configuration 3 (f = 2 needs n >= 11: n = 16)
run --aggregator median --nb-workers 8 --nb-decl-byz-workers 2                                                     # configuration 4
run --aggregator krum --nb-workers 8 --nb-decl-byz-workers 2 --nb-real-byz-workers 2 --attack flip --attack-args factor:-10 --no-baseline   # configuration 5
