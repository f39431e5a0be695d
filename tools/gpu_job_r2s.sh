#!/bin/bash
# round-2 GPU job S (final tree): gpu suite, default bench + reference arm, launch list, one ncu --set full capture of a round
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_broadcast.json 2> $O/bench_broadcast.err
echo "rc=$?" >> $O/bench_broadcast.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "rc=$?" >> $O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled --csv -c 3000 \
  --log-file $O/launches_broadcast.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_broadcast.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_round -s 720 -c 4 -f -o $O/prof_r2s \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_full.log 2>&1
tail -n 4 $O/pytest_gpu.log
for f in broadcast reference; do echo "== $f"; tail -n 1 $O/bench_$f.err; cut -c1-2600 $O/bench_$f.json; done
ls -la $O | head -20
