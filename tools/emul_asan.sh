#!/bin/bash
# Runs the emulated parity suites (kernel sources on the CPU SIMT emulator, tests/native/emul)
# under AddressSanitizer + UBSan: out-of-bounds accesses to "device" memory, which a GPU would
# silently absorb, abort the run here.  Usage: tools/emul_asan.sh [test files]
set -e
cd "$(dirname "$0")/.."
OUT=/tmp/libmaelstrom_b200_emul_asan.so
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -DMS_EMUL \
    -Itests/native/emul -Iinclude -x c++ maelstrom_b200/csrc/ms_kernels.cu -x c++ maelstrom_b200/csrc/ms_engine.cu \
    -x c++ tests/native/emul/simt.cpp -o "$OUT" -lpthread
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 MS_EMUL_LIB="$OUT" \
SUITES="tests/test_gpu_parity.py tests/test_emul_sharded.py tests/test_golden_fixtures.py tests/test_net_mirror.py
        tests/test_workload_gset.py tests/test_workload_services.py tests/test_workload_raft.py tests/test_workload_txn.py"
if [ $# -gt 0 ]; then SUITES="$*"; fi
python -m pytest $SUITES -q -m "not gpu"
