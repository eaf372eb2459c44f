#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_symeig.py -x -q --durations=5) > gpurun_out/b_symeig_test.log 2>&1
echo "rc=$?" >> gpurun_out/b_symeig_test.log
tail -15 gpurun_out/b_symeig_test.log
(timeout 300 python -m pytest tests/test_gpu_cpals.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -x -q -k "hooi or nvecs or tucker") > gpurun_out/b_hooi_test.log 2>&1
echo "rc=$?" >> gpurun_out/b_hooi_test.log
tail -5 gpurun_out/b_hooi_test.log
timeout 300 python tools/eig_perf.py > gpurun_out/b_eig_perf.log 2>&1
tail -12 gpurun_out/b_eig_perf.log
