#!/bin/bash
# round-2 GPU job T (final tree): ncu --set full of one HEAVY round of the broadcast bench (k_round launches 1604..1607 of
# the run = class 3, 2, 1, 0 of a round in the middle of a tick, located with profiles/r2s_launches_broadcast.csv), and the
# seeded hash-tree scenarios on the B200
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_round -s 1604 -c 4 -f -o $O/prof_r2t \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --journal-cap-log2 22 > $O/ncu_full.log 2>&1
MS_FUZZ_TREE_SEEDS=0:80 timeout 600 python -m pytest tests/test_txn_tree.py -m gpu -q > $O/tree_seeds_cuda.log 2>&1
echo "rc=$?" >> $O/tree_seeds_cuda.log
tail -n 3 $O/tree_seeds_cuda.log; tail -n 3 $O/ncu_full.log | cut -c1-300; ls -la $O
