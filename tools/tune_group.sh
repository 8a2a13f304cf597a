#!/bin/bash
# staged-kernel tile grouping (MXKV_B200_BULK_GROUP): sweep / BERT-base / ResNet-50 key sets at N=1, sweep at N=2
OUT=gpurun_out; mkdir -p $OUT
LIGHT="--no-e2e --no-cpu-baseline --no-sweep --no-secondary --no-parity"
pick() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']
print('%-10s G=%-3s ms/step %.4f kernel_ms %.4f frac %.3f' % ('$2', '$3', d['ms_per_step'], r['kernel_ms'], r['frac']))"; }
for G in 1 2 4 8 16 64; do
  export MXKV_B200_BULK_GROUP=$G
  timeout 120 python bench.py --steps 100 $LIGHT > $OUT/g_sweep_$G.json 2>/dev/null && pick $OUT/g_sweep_$G.json sweep $G
  timeout 120 python bench.py --workload bert --optimizer adam --steps 60 $LIGHT > $OUT/g_bert_$G.json 2>/dev/null && pick $OUT/g_bert_$G.json bert-adam $G
  timeout 120 python bench.py --workload resnet50 --optimizer sgd --steps 100 $LIGHT > $OUT/g_resnet_$G.json 2>/dev/null && pick $OUT/g_resnet_$G.json resnet-sgd $G
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 100 $LIGHT > $OUT/g_sweep2_$G.json 2>$OUT/g_sweep2_$G.err && pick $OUT/g_sweep2_$G.json sweep-N2 $G
done 2>&1 | tee $OUT/tune_bulk_group.txt
unset MXKV_B200_BULK_GROUP
grep -h "mxkv_b200\]" $OUT/g_sweep2_*.err | sort | uniq -c | tee -a $OUT/tune_bulk_group.txt
