#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_symeig.py -q --durations=3) > gpurun_out/d_symeig_test.log 2>&1
echo "rc=$?" >> gpurun_out/d_symeig_test.log
tail -25 gpurun_out/d_symeig_test.log
timeout 300 python tools/eig_perf.py > gpurun_out/d_eig_perf.log 2>&1
grep -E "symeig n=|^\{'n'|random|nvecs" gpurun_out/d_eig_perf.log | head -30
SKT_EIG_CLUSTER=8 timeout 300 python tools/eig_perf.py > gpurun_out/d_eig_perf_c8.log 2>&1
grep -E "symeig n=|^\{'n'" gpurun_out/d_eig_perf_c8.log | head -30
(timeout 300 python -m pytest tests/test_gpu_cpals.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_gpu_widerank.py -x -q -k "hooi or nvecs or tucker or wide or rank") > gpurun_out/d_hooi_test.log 2>&1
tail -5 gpurun_out/d_hooi_test.log
