#!/bin/bash
# round-2 GPU job D (N GPUs of one box): sharded parity tests, then the sharded bench with --verify
N=${1:-2}
mkdir -p gpurun_out/r2d_$N
O=gpurun_out/r2d_$N
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu > $O/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $O/pytest_sharded.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
tail -3 $O/pytest_sharded.log; tail -3 $O/bench.err | cut -c1-600; cut -c1-2500 $O/bench.json
