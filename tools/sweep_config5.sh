#!/bin/bash
# BASELINE configs[4]: synthetic KKT sweep over m at n = 1e6 on N GPUs. One bench.py line per m is appended to the output file
# (value = systems/s; roofline = condensation kernel in both modes; condensed_factor = the m x m Cholesky; timeline per rank).
# usage: tools/sweep_config5.sh N_GPUS OUT.jsonl [m ...]
N=${1:-1}; OUT=${2:-gpurun_out/sweep.jsonl}; shift 2
MS=${@:-256 512 1000 1024 2048 4000 4096 8192}
PORT=29600
for m in $MS; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --kkt-m $m --no-e2e --no-cpu >> "$OUT" 2>> "$OUT.err"
  else
    PORT=$((PORT+1))
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 3 --warmup 3 --kkt-m $m --no-e2e --no-cpu >> "$OUT" 2>> "$OUT.err"
  fi
  echo "m=$m N=$N rc=$?"
done
