#!/bin/bash
# compute-sanitizer over the GPU parity tests (run under gpurun from the repo root): memcheck over the lock-step / golden /
# property suites, racecheck over the crowded, ERVO and SFM cases.  Logs -> gpurun_out/ (copy into profiles/).
O=gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_properties_gpu.py -m gpu -q -k "not c4" > $O/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc $?"; tail -3 $O/r02_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -m gpu -q -k "crowded or ervo or c5_sfm or c3_orca" > $O/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc $?"; tail -3 $O/r02_sanitizer_racecheck.log
