#!/bin/bash
# round-2 GPU job C: gpu suite + compute-sanitizer over it, benches, launch lists, one ncu --set full capture
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
SEL='not cfg and not 4096 and not large_windows'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/memcheck.log \
  python -m pytest tests -m gpu -q -k "$SEL" > $O/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck_pytest.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $O/racecheck.log \
  python -m pytest tests -m gpu -q -k "$SEL" > $O/racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> $O/racecheck_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_broadcast.json 2> $O/bench_broadcast.err
echo "rc=$?" >> $O/bench_broadcast.err
for v in mb3 mb2; do
  MS_B200_LIB=$PWD/maelstrom_b200/libmaelstrom_b200_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_variant_$v.json 2> $O/bench_variant_$v.err
done
timeout 300 python bench.py --config broadcast-lat1 --values-per-tick 64 --steps 3 --warmup 3 --no-cpu > $O/bench_lat1_small.json 2> $O/bench_lat1_small.err
echo "rc=$?" >> $O/bench_lat1_small.err
timeout 600 python bench.py --config broadcast-lat1 --steps 6 --warmup 3 --no-cpu > $O/bench_lat1.json 2> $O/bench_lat1.err
echo "rc=$?" >> $O/bench_lat1.err
timeout 900 python bench.py --config gset16k --steps 3 --warmup 3 --no-cpu > $O/bench_gset16k.json 2> $O/bench_gset16k.err
echo "rc=$?" >> $O/bench_gset16k.err
timeout 600 python bench.py --config txn256k --steps 6 --warmup 3 --no-cpu > $O/bench_txn256k.json 2> $O/bench_txn256k.err
echo "rc=$?" >> $O/bench_txn256k.err
timeout 900 python bench.py --config raft64k --steps 6 --warmup 3 --no-cpu > $O/bench_raft64k.json 2> $O/bench_raft64k.err
echo "rc=$?" >> $O/bench_raft64k.err
# launch lists (cold, serialised: shares only)
for c in gset16k txn256k raft64k; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled --csv -c 6000 \
    --log-file $O/launches_$c.csv python bench.py --config $c --steps 1 --warmup 3 --no-cpu --no-e2e > $O/ncu_$c.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled --csv -c 3000 \
  --log-file $O/launches_broadcast.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_broadcast.log 2>&1
# one whole round of the broadcast bench, full sections + source counters (4 size-class launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_round -s 720 -c 4 -f -o $O/prof_r2c \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_full.log 2>&1
tail -3 $O/pytest_gpu.log $O/memcheck_pytest.log $O/racecheck_pytest.log
tail -c 300 $O/memcheck.log $O/racecheck.log
for f in broadcast variant_mb3 variant_mb2 lat1_small lat1 gset16k txn256k raft64k; do echo "== $f"; tail -2 $O/bench_$f.err | cut -c1-700; cut -c1-400 $O/bench_$f.json; done
ls -la $O
