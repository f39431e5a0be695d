#!/bin/bash
# round-2 GPU job V: the final library (phase timing compiled out): gpu suite + default bench
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_broadcast.json 2> $O/bench_broadcast.err
echo "rc=$?" >> $O/bench_broadcast.err
tail -n 3 $O/pytest_gpu.log; tail -n 1 $O/bench_broadcast.err; cut -c1-2300 $O/bench_broadcast.json
