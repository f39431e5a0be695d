#!/bin/bash
# round-2 GPU job G: kernel tuning A/B -- the broadcast bench (device arm only) on the current library and on
# variant builds, then the fast gpu parity suites on the current library
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
for v in "" _v3 _mb3; do
  MS_B200_LIB=$PWD/maelstrom_b200/libmaelstrom_b200$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench$v.json 2> $O/bench$v.err
  echo "rc=$?" >> $O/bench$v.err
done
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
for v in "" _v3 _mb3; do echo "== $v"; tail -1 $O/bench$v.err; python - <<PY
import json
try:
    d = json.load(open("$O/bench$v.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["config"]["fallback_sorts"])
except Exception as e:
    print("no json", e)
PY
done
tail -3 $O/pytest_gpu.log
