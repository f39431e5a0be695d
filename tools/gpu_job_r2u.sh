#!/bin/bash
# round-2 GPU job U: the hash-tree scenarios again after the lock-queue change (final library)
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
MS_FUZZ_TREE_SEEDS=0:150 timeout 600 python -m pytest tests/test_txn_tree.py -m gpu -q > $O/tree_seeds_cuda.log 2>&1
echo "rc=$?" >> $O/tree_seeds_cuda.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "rc=$?" >> $O/smoke.log
tail -n 3 $O/tree_seeds_cuda.log $O/smoke.log
