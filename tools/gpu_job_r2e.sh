#!/bin/bash
# round-2 GPU job E: gpu suite, benches of every config (N=1), launch list + one ncu --set full capture of a broadcast round
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_broadcast.json 2> $O/bench_broadcast.err
echo "rc=$?" >> $O/bench_broadcast.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "rc=$?" >> $O/bench_reference.err
timeout 600 python bench.py --config broadcast-lat1 --steps 6 --warmup 3 --no-cpu > $O/bench_lat1.json 2> $O/bench_lat1.err
echo "rc=$?" >> $O/bench_lat1.err
timeout 900 python bench.py --config gset16k --steps 3 --warmup 3 --no-cpu > $O/bench_gset16k.json 2> $O/bench_gset16k.err
echo "rc=$?" >> $O/bench_gset16k.err
timeout 600 python bench.py --config txn256k --steps 6 --warmup 3 --no-cpu > $O/bench_txn256k.json 2> $O/bench_txn256k.err
echo "rc=$?" >> $O/bench_txn256k.err
timeout 900 python bench.py --config raft64k --steps 6 --warmup 3 --no-cpu > $O/bench_raft64k.json 2> $O/bench_raft64k.err
echo "rc=$?" >> $O/bench_raft64k.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled --csv -c 3000 \
  --log-file $O/launches_broadcast.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_broadcast.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_round -s 720 -c 4 -f -o $O/prof_r2e \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_full.log 2>&1
tail -3 $O/pytest_gpu.log
for f in broadcast reference lat1 gset16k txn256k raft64k; do echo "== $f"; tail -2 $O/bench_$f.err | cut -c1-500; cut -c1-1500 $O/bench_$f.json; done
ls -la $O
