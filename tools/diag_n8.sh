#!/bin/bash
# r02: why does the 8-GPU bench line (1.07 ms) differ from tools/tune_nvls.py (0.767 ms) on the same kernel?  One variable at a time.
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1 MXKV_B200_SPIN_TIMEOUT_S=20
LIGHT="--no-e2e --no-cpu-baseline --no-sweep --no-secondary --no-parity"
tr8() { echo python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)); }
pick() { python -c "
import json
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']
print('%-34s ms/step %.4f kernel_ms %.4f  %s' % ('$2', d['ms_per_step'], r['kernel_ms'], d.get('exchange')))" 2>&1 | tail -1; }
{
timeout 150 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider -k "sp_" 2>&1 | tail -3
TUNE_QUICK=1 TUNE_GRIDS=48,0 TUNE_OPTS=sgd timeout 120 $(tr8) tools/tune_nvls.py 2>/dev/null | grep -E "nvls|p2p bulk|multicast memory|nccl all_reduce flat"
timeout 120 $(tr8) bench.py --gpus 8 --steps 200 $LIGHT > $OUT/d_a.json 2>$OUT/d_a.err; pick $OUT/d_a.json "bench default (sampler+nvml)"
timeout 120 $(tr8) bench.py --gpus 8 --steps 200 --no-clocks $LIGHT > $OUT/d_b.json 2>$OUT/d_b.err; pick $OUT/d_b.json "bench --no-clocks"
timeout 120 $(tr8) bench.py --gpus 8 --steps 200 --no-clocks --no-nvml $LIGHT > $OUT/d_c.json 2>$OUT/d_c.err; pick $OUT/d_c.json "bench --no-clocks --no-nvml"
MXKV_B200_PLAN=0 timeout 120 $(tr8) bench.py --gpus 8 --steps 200 --no-clocks --no-nvml $LIGHT > $OUT/d_d.json 2>$OUT/d_d.err; pick $OUT/d_d.json "same, MXKV_B200_PLAN=0"
} 2>&1 | tee $OUT/diag_n8.txt
