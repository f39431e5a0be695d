#!/bin/bash
# round-2 GPU job A: baseline of the round-1 tree on hardware + compute-sanitizer over the golden cases
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
SEL='golden or doc_message or doc_counts or smoke_entry or flood_counts'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2a/memcheck.log \
  python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r2a/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2a/memcheck_pytest.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r2a/racecheck.log \
  python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r2a/racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2a/racecheck_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
echo "bench rc=$?" >> gpurun_out/r2a/bench.err
tail -3 gpurun_out/r2a/pytest_gpu.log gpurun_out/r2a/memcheck_pytest.log gpurun_out/r2a/racecheck_pytest.log
tail -c 600 gpurun_out/r2a/memcheck.log gpurun_out/r2a/racecheck.log
cat gpurun_out/r2a/bench.json
