#!/bin/bash
# Race check on the CPU: the kernel + engine sources compiled with ThreadSanitizer on the SIMT emulator,
# every emulated thread a TSan fiber (tests/native/emul/simt.cpp, MS_TSAN), driven through the C ABI by
# tests/native/emul/tsan_driver.cpp.  Reports = two threads of one CTA touching the same location with no
# __syncthreads() / warp collective in between.  Usage: tools/emul_tsan.sh
set -e
cd "$(dirname "$0")/.."
B=/tmp/ms_tsan_build
mkdir -p $B
F="-O1 -g -std=c++17 -fPIC -DMS_EMUL -DMS_TSAN -Itests/native/emul -Iinclude"
g++ $F -c tests/native/emul/simt.cpp -o $B/simt.o                                   # NOT instrumented
g++ $F -fsanitize=thread -x c++ -c maelstrom_b200/csrc/ms_kernels.cu -o $B/kernels.o
g++ $F -fsanitize=thread -x c++ -c maelstrom_b200/csrc/ms_engine.cu -o $B/engine.o
g++ $F -c tests/native/emul/tsan_driver.cpp -o $B/driver.o
g++ -fsanitize=thread $B/driver.o $B/kernels.o $B/engine.o $B/simt.o -o $B/tsan_driver -lpthread
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4 suppressions=tests/native/emul/tsan.supp" $B/tsan_driver "$@"
