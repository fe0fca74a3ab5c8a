#!/bin/bash
# 70B-shape layer stack (SURVEY 8e / BASELINE config 5) on N GPUs of one box: N-sharded linears, output all-gather
# fused into the GEMV epilogue.  usage: tools/scale_70b.sh N [extra bench args]
N=$1; shift
OUT=gpurun_out/scale70b_n${N}.jsonl
: > $OUT
run() {
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" 2>>gpurun_out/scale70b_n${N}.err | tee -a $OUT
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>>gpurun_out/scale70b_n${N}.err | tee -a $OUT; fi
}
run --workload llama3-70b --steps 20 --warmup 5 --no-cpu-baseline "$@"
if [ "$N" != "1" ]; then
  run --workload llama3-70b --steps 20 --warmup 5 --no-cpu-baseline --collective nccl "$@"
  run --workload llama2-7b --steps 20 --warmup 5 --no-cpu-baseline "$@"
fi
