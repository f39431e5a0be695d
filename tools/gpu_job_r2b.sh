#!/bin/bash
# round-2 GPU job B: new tree -- full gpu test-suite, bench of every config at N=1
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_broadcast.json 2> $O/bench_broadcast.err
echo "rc=$?" >> $O/bench_broadcast.err
timeout 600 python bench.py --config broadcast-lat1 --steps 6 --warmup 3 --no-cpu > $O/bench_lat1.json 2> $O/bench_lat1.err
echo "rc=$?" >> $O/bench_lat1.err
timeout 900 python bench.py --config gset16k --steps 3 --warmup 3 --no-cpu > $O/bench_gset16k.json 2> $O/bench_gset16k.err
echo "rc=$?" >> $O/bench_gset16k.err
timeout 600 python bench.py --config txn256k --steps 6 --warmup 3 --no-cpu > $O/bench_txn256k.json 2> $O/bench_txn256k.err
echo "rc=$?" >> $O/bench_txn256k.err
timeout 900 python bench.py --config raft64k --steps 6 --warmup 3 --no-cpu > $O/bench_raft64k.json 2> $O/bench_raft64k.err
echo "rc=$?" >> $O/bench_raft64k.err
tail -4 $O/pytest_gpu.log
for f in broadcast lat1 gset16k txn256k raft64k; do echo "== $f"; tail -3 $O/bench_$f.err; cut -c1-600 $O/bench_$f.json; done
