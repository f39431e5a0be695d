#!/bin/bash
# round-2 GPU job I: memcheck over the five families again (new library: stall report in the error text), then initcheck
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
SEL='not cfg and not 4096 and not large_windows and not heavy and not fuzz'
FILES="tests/test_gpu_parity.py tests/test_workload_gset.py tests/test_workload_services.py tests/test_workload_raft.py tests/test_workload_txn.py tests/test_gen_clients.py tests/test_journal_stream.py"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/memcheck.log \
  python -m pytest $FILES -m gpu -q -x -k "$SEL" > $O/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck_pytest.log
timeout 700 compute-sanitizer --tool initcheck --error-exitcode 7 --log-file $O/initcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $O/initcheck_pytest.log 2>&1
echo "initcheck rc=$?" >> $O/initcheck_pytest.log
grep -h "^E  " $O/memcheck_pytest.log | cut -c1-900
tail -n 3 $O/memcheck_pytest.log $O/initcheck_pytest.log
grep -c "Uninitialized" $O/initcheck.log; grep -m6 -A4 "Uninitialized" $O/initcheck.log | cut -c1-220
tail -n 2 $O/memcheck.log
