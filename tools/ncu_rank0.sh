#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs under ncu (NVLink / DRAM byte counters and the duration of the dense
# kernels, ONE pass -- a kernel that rendezvous with its peers cannot be replayed), the other ranks run plain.
#   python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node N ... tools/ncu_rank0.sh bench.py <args>
# Output: gpurun_out/ncu_nvl_n$WORLD_SIZE.csv.  A number printed by this run is never a bench value.
export MXKV_B200_SPIN_TIMEOUT_S=${MXKV_B200_SPIN_TIMEOUT_S:-20}
M=${NCU_METRICS:-gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvltx__bytes_data_protocol.sum,dram__bytes_read.sum,dram__bytes_write.sum}
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics $M --clock-control none --replay-mode kernel -k regex:${NCU_KERNEL:-kv_dense} -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-3} \
       --csv --log-file gpurun_out/ncu_nvl_n${WORLD_SIZE:-1}${NCU_TAG:-}.csv python "$@"
else
  exec python "$@"
fi
