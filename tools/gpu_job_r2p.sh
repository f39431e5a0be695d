#!/bin/bash
# round-2 GPU job P (N GPUs): the sharded bench as the driver's scaling run launches it (verify is on by default for N > 1)
N=${1:-8}
mkdir -p gpurun_out/r2p_$N
O=gpurun_out/r2p_$N
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 6 --warmup 3 > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
tail -n 3 $O/bench.err | cut -c1-600; cut -c1-3000 $O/bench.json
