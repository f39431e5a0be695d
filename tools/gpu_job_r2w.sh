#!/bin/bash
# round-2 GPU job W: first timing of the hash-tree txn-list-append node program (bench.py --config txntree)
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 500 python bench.py --config txntree --steps 6 --warmup 3 --no-cpu > $O/bench_txntree.json 2> $O/bench_txntree.err
echo "rc=$?" >> $O/bench_txntree.err
tail -n 3 $O/bench_txntree.err | cut -c1-600; cut -c1-2000 $O/bench_txntree.json
