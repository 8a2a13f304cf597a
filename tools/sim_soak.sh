#!/bin/bash
# The `-m gpu` parity tests on the simulated runtime (tests/sim: the dense kernels run from their own source on CPU
# threads) under tuning knobs the default CPU suite does not visit: grouped tile walks, small / deep rings of the staged
# kernel, other chunk and block sizes, every key sharded, plus extra random walks.  No GPU needed; ~3 minutes on 8 cores.
#   bash tools/sim_soak.sh          (builds tests/sim/_build first if needed)
cd "$(dirname "$0")/.."
python -c "import sys; sys.path.insert(0, 'tests/sim'); import build_sim; build_sim.build()" || exit 1
export MXKV_SIM=1 MXKV_B200_LIBRARY_PATH=tests/sim/_build/libmxkv_b200_sim.so
FAIL=0
run() {   # run <devices> [VAR=value ...] <pytest arguments>
  local D=$1; shift
  echo "== devices=$D $*"
  env MXKV_SIM_DEVICES=$D "$@" 2>&1 | grep -v "try backtracking" | tail -2
  [ "${PIPESTATUS[0]}" = "0" ] || FAIL=1
}
P="timeout 3000 python -m pytest -q -m gpu -p no:cacheprovider -n 6 -x"
run 1 MXKV_B200_BULK_GROUP=4 $P tests/test_gpu_dense.py
run 1 MXKV_B200_BULK_GROUP=3 MXKV_B200_BULK_TILE=256 MXKV_B200_BULK_STAGES=3 $P tests/test_gpu_dense.py
run 1 MXKV_B200_BULK_TILE=512 MXKV_B200_BULK_STAGES=8 MXKV_B200_CHUNK=1024 $P tests/test_gpu_dense.py tests/test_gpu_y_semantics.py
run 1 MXKV_B200_BULK=2 MXKV_B200_THREADS=128 $P tests/test_gpu_dense.py tests/test_gpu_reference_kats.py
# the same dense kernels on the OS-thread engine: a block's threads really run at the same time
run 1 MXKV_SIM_ENGINE=threads $P tests/test_gpu_dense.py
run 4 MXKV_SIM_ENGINE=threads MXKV_B200_BULK=2 $P tests/test_gpu_multi.py -k "not one_process_per_gpu"
run 4 MXKV_B200_BULK=2 MXKV_B200_BULK_GROUP=4 MXKV_B200_BULK_TILE=256 $P tests/test_gpu_multi.py tests/test_gpu_y_placement.py -k "not one_process_per_gpu"
run 4 MXKV_B200_BULK=0 MXKV_B200_CHUNK=256 $P tests/test_gpu_multi.py tests/test_gpu_y_placement.py -k "not one_process_per_gpu"
run 3 MXKV_FUZZ_SEEDS=60 $P tests/test_gpu_y_placement.py -k randomized
run 8 MXKV_FUZZ_SEEDS=60 MXKV_B200_BULK=2 MXKV_B200_TWOSHOT_BYTES=4096 $P tests/test_gpu_y_placement.py -k randomized
for D in 3 5 8; do run $D MXKV_FUZZ_SEEDS=150 MXKV_B200_CHUNK=512 $P tests/test_gpu_zzz_tree.py; done
# the multicast arena's agree-or-fall-back protocol with a driver call failing at every stage (three processes each)
echo "== multicast arena: every failure stage"
env -u MXKV_SIM -u MXKV_B200_LIBRARY_PATH MXKV_FUZZ_SEEDS=1 timeout 1500 python -m pytest -q -p no:cacheprovider tests/test_sim_host_logic.py -k arena_and_its_fallbacks 2>&1 | tail -2
[ "${PIPESTATUS[0]}" = "0" ] || FAIL=1
[ $FAIL = 0 ] && echo SOAK_OK || echo SOAK_FAILED
exit $FAIL
