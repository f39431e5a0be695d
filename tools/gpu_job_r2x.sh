#!/bin/bash
# round-2 GPU job X: the final library after the promise time-outs: full gpu suite, hash-tree seeds, smoke
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
MS_FUZZ_TREE_SEEDS=0:60 timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "rc=$?" >> $O/smoke.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
tail -n 3 $O/pytest_gpu.log $O/smoke.log; tail -n 1 $O/bench.err; cut -c1-400 $O/bench.json
