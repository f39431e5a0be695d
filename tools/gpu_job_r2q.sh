#!/bin/bash
# round-2 GPU job Q (2 GPUs): k_glue on hardware -- sharded parity tests, then the sharded bench with and without it
N=${1:-2}
mkdir -p gpurun_out/r2q_$N
O=gpurun_out/r2q_$N
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu > $O/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $O/pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > $O/bench_glue.json 2> $O/bench_glue.err
echo "rc=$?" >> $O/bench_glue.err
MS_NO_GLUE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-verify > $O/bench_noglue.json 2> $O/bench_noglue.err
echo "rc=$?" >> $O/bench_noglue.err
tail -n 3 $O/pytest_sharded.log
for v in glue noglue; do echo "== $v"; tail -n 1 $O/bench_$v.err; python - <<PY
import json
try:
    d = json.load(open("$O/bench_$v.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_us"], d["gpu_launches"], d.get("parity_digest_ok"), d["clocks"])
except Exception as e:
    print("no json", e)
PY
done
