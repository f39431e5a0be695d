#!/bin/bash
# round-2 GPU job Y: the other workload classes on the final library (device arm + e2e, no CPU leg)
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
timeout 200 python bench.py --config broadcast-lat1 --steps 6 --warmup 3 --no-cpu > $O/bench_lat1.json 2> $O/bench_lat1.err
timeout 300 python bench.py --config gset16k --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_gset16k.json 2> $O/bench_gset16k.err
timeout 200 python bench.py --config txn256k --steps 6 --warmup 3 --no-cpu > $O/bench_txn256k.json 2> $O/bench_txn256k.err
timeout 200 python bench.py --config raft64k --steps 6 --warmup 3 --no-cpu > $O/bench_raft64k.json 2> $O/bench_raft64k.err
for f in lat1 gset16k txn256k raft64k; do echo "== $f"; tail -n 1 $O/bench_$f.err | cut -c1-200; cut -c1-330 $O/bench_$f.json; done
