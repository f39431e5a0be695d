#!/bin/bash
# round-2 GPU job K: race fix validation -- full gpu suite, memcheck over the five families, default bench (all legs)
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
SEL='not cfg and not 4096 and not large_windows and not heavy and not fuzz'
FILES="tests/test_gpu_parity.py tests/test_workload_gset.py tests/test_workload_services.py tests/test_workload_raft.py tests/test_workload_txn.py tests/test_gen_clients.py tests/test_journal_stream.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/memcheck.log \
  python -m pytest $FILES -m gpu -q -k "$SEL" > $O/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_broadcast.json 2> $O/bench_broadcast.err
echo "rc=$?" >> $O/bench_broadcast.err
tail -n 3 $O/pytest_gpu.log $O/memcheck_pytest.log; tail -n 2 $O/memcheck.log; tail -n 2 $O/bench_broadcast.err; cut -c1-2400 $O/bench_broadcast.json
