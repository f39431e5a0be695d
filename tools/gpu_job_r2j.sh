#!/bin/bash
# round-2 GPU job J: (1) the memcheck stall with one round per batch and the device state after every batch;
# (2) class-width A/B and per-phase ticket latency of the broadcast bench
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
SEL='not cfg and not 4096 and not large_windows and not heavy and not fuzz'
MS_DEBUG_STALL=1 timeout 600 compute-sanitizer --tool memcheck --log-file $O/memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $O/memcheck_pytest.log 2> $O/memcheck_stderr.log
echo "memcheck rc=$?" >> $O/memcheck_pytest.log
for v in "" _c0w32 _c0w32c1w64; do
  MS_B200_LIB=$PWD/maelstrom_b200/libmaelstrom_b200$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench$v.json 2> $O/bench$v.err
  echo "rc=$?" >> $O/bench$v.err
done
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --phase-cycles > $O/phase.json 2> $O/phase.txt
grep -h "^E  " $O/memcheck_pytest.log | cut -c1-600
tail -n 2 $O/memcheck_pytest.log
grep "MS_DEBUG_STALL" $O/memcheck_stderr.log $O/memcheck_pytest.log | tail -n 12 | cut -c1-500
for v in "" _c0w32 _c0w32c1w64; do echo "== $v"; tail -n 1 $O/bench$v.err; python - <<PY
import json
try:
    d = json.load(open("$O/bench$v.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["gpu_launches"])
except Exception as e:
    print("no json", e)
PY
done
cat $O/phase.txt | tail -n 45
