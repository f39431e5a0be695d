#!/bin/bash
# round-2 GPU job ZZ: rounds per streamed batch, A/B on the e2e leg
mkdir -p gpurun_out/r2zz
O=gpurun_out/r2zz
for b in 32 8; do
  MS_STREAM_BATCH_ROUNDS=$b timeout 100 python bench.py --steps 4 --warmup 3 --no-cpu > $O/bench_b$b.json 2> $O/bench_b$b.err
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_b$b.json")); print("batch $b", d["value"], d["e2e"]["value"])
except Exception as e:
    print("batch $b no json", e)
PY
done
