#!/bin/bash
# One gpurun call's worth of validation + measurement for N GPUs, every step under its own timeout, everything
# written to gpurun_out/ (merged back by gpurun).  Usage:
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh 1'
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_round.sh 2'
#   gpurun --gpus 8 --timeout 1800 -- 'bash tools/gpu_round.sh 8 [tune|tree]'
# Order = priority: parity first (all failures listed, not just the first), then the bench lines, then profiles.
N=${1:-1}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
LIGHT="--no-e2e --no-cpu-baseline --no-sweep --no-secondary --no-parity"
run() {   # run <seconds> <logfile> <command...>
  local t=$1 log=$2; shift 2
  echo "== $* (limit ${t}s)" | tee -a $OUT/round_n$N.log
  local t0=$(date +%s)
  timeout $t "$@" > $OUT/$log 2>$OUT/$log.err
  echo "   exit $? after $(( $(date +%s) - t0 ))s -> $OUT/$log" | tee -a $OUT/round_n$N.log
}
# (a plain word list, not a shell function: `timeout` execs its command)
torchrun_() { echo python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)); }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpus_n$N.csv 2>&1
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1

if [ "$N" = "1" ]; then
  run 1500 pytest_n1.log python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=12
  tail -n 25 $OUT/pytest_n1.log
  run 200 smoke_n1.log python __graft_entry__.py smoke
  run 700 bench_n1.json python bench.py --gpus 1
  run 300 launches_n1.csv ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv \
      python bench.py --steps 20 --warmup 3 $LIGHT
  run 400 ncu_full_n1.log ncu --set full --clock-control none --import-source on -k regex:kv_dense_bulk -s 4 -c 2 \
      -f -o $OUT/r02_n1_bulk python bench.py --steps 6 --warmup 3 $LIGHT
  run 300 bench_bert_adam_n1.json python bench.py --workload bert --optimizer adam --steps 40 $LIGHT
  run 300 bench_bert_lamb_n1.json python bench.py --workload bert --optimizer lamb --steps 40 $LIGHT
  run 300 bench_resnet_sgd_n1.json python bench.py --workload resnet50 --optimizer sgd --steps 100 $LIGHT
  run 300 bench_ref_n1.json python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
elif [ "$MODE" = "tree" ]; then
  # MXNET_KVSTORE_USETREE=1: written after round 2's GPU minutes were spent -- first hardware run.
  # Parity first (single process 3..N GPUs, then one process per GPU), then the bench line in tree order (its parity
  # object follows the variable) next to the default order on the same box.   gpurun --gpus 4 -- 'bash tools/gpu_round.sh 4 tree'
  run 600 pytest_tree_n$N.log python -m pytest tests/test_gpu_zzz_tree.py -m gpu -q --maxfail=25 -p no:cacheprovider --durations=8
  tail -n 12 $OUT/pytest_tree_n$N.log
  run 600 pytest_mp_n$N.log python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -k "one_process_per_gpu and not hierarchy"
  MXNET_KVSTORE_USETREE=1 MXNET_KVSTORE_LOGTREE=1 run 300 bench_tree_n$N.json $(torchrun_ $N) bench.py --gpus $N --steps 200 --no-e2e --no-cpu-baseline --no-sweep --no-secondary
  MXKV_B200_NVLS=0 run 300 bench_plain_n$N.json $(torchrun_ $N) bench.py --gpus $N --steps 200 --no-e2e --no-cpu-baseline --no-sweep --no-secondary
elif [ "$MODE" = "tune" ]; then
  # kernel tuning only (8-GPU minutes are charged 8x): multicast kernel knobs, peer kernels, NCCL, then N=4 on the same box
  run 600 tune_nvls_n$N.txt $(torchrun_ $N) tools/tune_nvls.py
  if [ "$N" = "8" ]; then
    TUNE_GRIDS=0,148,64,32 TUNE_OPTS=sgd run 300 tune_nvls_n4.txt $(torchrun_ 4) tools/tune_nvls.py
  fi
elif [ "$N" = "8" ]; then
  # 8-GPU minutes are charged 8x: every multi-GPU test at HEAD (the single-GPU forced-variant sweeps ran on the
  # 1- and 2-GPU boxes), the bench line, the hierarchy on hardware, engine-owned vs torch-owned multicast memory
  run 900 pytest_n8.log python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=8 -k "not forced_variant"
  tail -n 14 $OUT/pytest_n8.log
  run 500 bench_n8.json $(torchrun_ 8) bench.py --gpus 8
  run 200 bench_ref_n8.json $(torchrun_ 8) bench.py --impl reference --gpus 8 --steps 3 --warmup 1
  run 200 bench_n8_hier_2x4.json $(torchrun_ 8) bench.py --gpus 8 --local-world 4 --steps 100 $LIGHT
  MXKV_B200_ARENA_VMM=0 run 200 bench_n8_torch_multicast.json $(torchrun_ 8) bench.py --gpus 8 --steps 200 $LIGHT
  run 200 bench_bert_adam_n8.json $(torchrun_ 8) bench.py --gpus 8 --workload bert --optimizer adam --steps 40 $LIGHT
else
  run 1500 pytest_n$N.log python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=12
  tail -n 25 $OUT/pytest_n$N.log
  run 700 bench_n$N.json $(torchrun_ $N) bench.py --gpus $N
  run 300 bench_ref_n$N.json $(torchrun_ $N) bench.py --impl reference --gpus $N --steps 3 --warmup 1
  run 300 bench_bert_adam_n$N.json $(torchrun_ $N) bench.py --gpus $N --workload bert --optimizer adam --steps 40 $LIGHT
  if [ "$N" = "2" ]; then      # the box has the time: the single-GPU lines of this commit as well
    run 400 bench_n1.json python bench.py --gpus 1
    run 200 bench_bert_adam_n1.json python bench.py --workload bert --optimizer adam --steps 40 $LIGHT
    run 200 bench_resnet_sgd_n1.json python bench.py --workload resnet50 --optimizer sgd --steps 100 $LIGHT
    run 200 launches_bert_n1.csv ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 20 --csv \
        python bench.py --workload bert --optimizer adam --steps 12 --warmup 3 $LIGHT
    MXKV_B200_ARENA_VMM=0 run 200 bench_n2_ipc_arena.json $(torchrun_ 2) bench.py --gpus 2 --steps 300 $LIGHT
    run 200 bench_n2_vmm_arena.json $(torchrun_ 2) bench.py --gpus 2 --steps 300 $LIGHT
  fi
  if [ "$N" = "4" ]; then
    run 400 bench_n4_hier_2x2.json $(torchrun_ 4) bench.py --gpus 4 --local-world 2 --steps 100 $LIGHT
  fi
fi
grep -h '"metric"' $OUT/bench*_n$N*.json 2>/dev/null | cut -c1-600
