#!/bin/bash
# One gpurun call's worth of validation + measurement for N GPUs, every step under its own timeout, everything
# written to gpurun_out/ (merged back by gpurun).  Usage:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh 1'
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_round.sh 2'
#   gpurun --gpus 8 --timeout 1800 -- 'bash tools/gpu_round.sh 8'
# Order = priority: parity first (all failures listed, not just the first), then the bench lines, then profiles.
N=${1:-1}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run() {   # run <seconds> <logfile> <command...>
  local t=$1 log=$2; shift 2
  echo "== $* (limit ${t}s)" | tee -a $OUT/round_n$N.log
  timeout $t "$@" > $OUT/$log 2>&1
  echo "   exit $? -> $OUT/$log" | tee -a $OUT/round_n$N.log
}
torchrun_() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 "$@"; }

run 900 pytest_n$N.log python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider
tail -n 15 $OUT/pytest_n$N.log
run 120 smoke_n$N.log python __graft_entry__.py smoke

if [ "$N" = "1" ]; then
  run 400 bench_n1.json python bench.py --gpus 1
  for seg in 1048576 2097152 8388608; do      # e2e pipeline segment (default 4 Mi elements): fill/drain vs launch count
    MXKV_B200_HOST_SEG_ELEMS=$seg run 200 bench_n1_seg$seg.json python bench.py --gpus 1 --steps 40 --no-cpu-baseline
  done
  run 300 bench_bert_adam_n1.json python bench.py --workload bert --optimizer adam --steps 40 --no-e2e --no-cpu-baseline
  run 300 bench_bert_lamb_n1.json python bench.py --workload bert --optimizer lamb --steps 40 --no-e2e --no-cpu-baseline
  run 300 launches_n1.txt ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 80 --csv \
      python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline
else
  run 500 bench_n$N.json torchrun_ bench.py --gpus $N
  run 300 bench_bert_adam_n$N.json torchrun_ bench.py --gpus $N --workload bert --optimizer adam --steps 40 --no-e2e --no-cpu-baseline
  if [ "$N" = "8" ]; then
    for u in 4 8; do
      MXKV_B200_NVLS_U=$u run 300 bench_n8_nvls_u$u.json torchrun_ bench.py --gpus 8 --steps 100 --no-e2e --no-cpu-baseline
    done
    MXKV_B200_NVLS_U=8 run 300 bench_n8_allreduce_u8.json torchrun_ bench.py --gpus 8 --steps 100 --optimizer none --no-e2e --no-cpu-baseline
    run 400 bench_n8_hier_2x4.json torchrun_ bench.py --gpus 8 --local-world 4 --steps 100 --no-e2e --no-cpu-baseline
  fi
  if [ "$N" = "4" ]; then
    run 400 bench_n4_hier_2x2.json torchrun_ bench.py --gpus 4 --local-world 2 --steps 100 --no-e2e --no-cpu-baseline
  fi
fi
grep -h '"metric"' $OUT/bench*_n$N*.json 2>/dev/null | cut -c1-400
