#!/bin/bash
# round-2 GPU job N: the differential fuzzer on the product library (B200) instead of the emulator
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
MS_FUZZ_BACKEND=cuda MS_FUZZ_SEEDS=0:1500 MS_FUZZ_HEAVY_SEEDS=0:120 MS_FUZZ_SHARDED_SEEDS=0:0 MS_FUZZ_RAFT_SEEDS=0:40 \
  timeout 1500 python -m pytest tests/test_fuzz_parity.py -q -x -p no:cacheprovider > $O/fuzz_cuda.log 2>&1
echo "fuzz rc=$?" >> $O/fuzz_cuda.log
tail -n 6 $O/fuzz_cuda.log
