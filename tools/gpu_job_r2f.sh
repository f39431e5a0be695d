#!/bin/bash
# round-2 GPU job F: the re-sized scale tests, then compute-sanitizer (memcheck, racecheck) over the [cuda] parity cases
# of all five node-program families (echo/broadcast, g-set, services, Raft, txn-list-append) at golden-case sizes
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_scale_configs.py -m gpu -q --durations=5 > $O/pytest_fixed.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fixed.log
SEL='not cfg and not 4096 and not large_windows and not heavy and not fuzz'
FILES="tests/test_gpu_parity.py tests/test_workload_gset.py tests/test_workload_services.py tests/test_workload_raft.py tests/test_workload_txn.py tests/test_gen_clients.py tests/test_journal_stream.py"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/memcheck.log \
  python -m pytest $FILES -m gpu -q -k "$SEL" > $O/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck_pytest.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $O/racecheck.log \
  python -m pytest $FILES -m gpu -q -k "$SEL" > $O/racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> $O/racecheck_pytest.log
tail -4 $O/pytest_fixed.log $O/memcheck_pytest.log $O/racecheck_pytest.log
tail -c 400 $O/memcheck.log $O/racecheck.log
