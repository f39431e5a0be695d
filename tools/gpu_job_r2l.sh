#!/bin/bash
# round-2 GPU job L: MS_JFMT_4 on hardware -- stream tests, default bench (e2e at 4 B/event), e2e at 8 B/event for comparison
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 600 python -m pytest tests/test_journal_stream.py tests/test_sass_fingerprint.py -m gpu -q > $O/pytest_stream.log 2>&1
echo "pytest rc=$?" >> $O/pytest_stream.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_fmt4.json 2> $O/bench_fmt4.err
echo "rc=$?" >> $O/bench_fmt4.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --stream-format 8 > $O/bench_fmt8.json 2> $O/bench_fmt8.err
echo "rc=$?" >> $O/bench_fmt8.err
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu --touch > $O/bench_fmt4_touch.json 2> $O/bench_fmt4_touch.err
echo "rc=$?" >> $O/bench_fmt4_touch.err
tail -n 3 $O/pytest_stream.log
for v in fmt4 fmt8 fmt4_touch; do echo "== $v"; tail -n 2 $O/bench_$v.err | cut -c1-300; python - <<PY
import json
try:
    d = json.load(open("$O/bench_$v.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"])
except Exception as e:
    print("no json", e)
PY
done
