#!/bin/bash
# round-2 GPU job H: the two tests that stop advancing under memcheck, one at a time with the stall report, then initcheck
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
T1='tests/test_gpu_parity.py::test_sim_clients_and_block_path_counters'
T2='tests/test_workload_raft.py::test_election_replication_proxy_and_errors'
timeout 300 python -m pytest "$T1" "$T2" -m gpu -q > $O/plain.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --log-file $O/memcheck_1.log python -m pytest "$T1" -m gpu -q > $O/memcheck_1_pytest.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --log-file $O/memcheck_2.log python -m pytest "$T2" -m gpu -q > $O/memcheck_2_pytest.log 2>&1
timeout 500 compute-sanitizer --tool initcheck --track-unused-memory no --log-file $O/initcheck_1.log python -m pytest "$T1" -m gpu -q > $O/initcheck_1_pytest.log 2>&1
tail -n 3 $O/plain.log
grep -h "^E  " $O/memcheck_1_pytest.log $O/memcheck_2_pytest.log | cut -c1-700
tail -n 3 $O/memcheck_1_pytest.log $O/memcheck_2_pytest.log $O/initcheck_1_pytest.log
grep -c "Uninitialized" $O/initcheck_1.log; grep -m8 -A3 "Uninitialized" $O/initcheck_1.log | cut -c1-200
