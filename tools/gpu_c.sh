#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_symeig.py -q --durations=5) > gpurun_out/c_symeig_test.log 2>&1
echo "rc=$?" >> gpurun_out/c_symeig_test.log
tail -25 gpurun_out/c_symeig_test.log
(timeout 300 python -m pytest tests/test_gpu_widerank.py -q) > gpurun_out/c_wide_test.log 2>&1
echo "rc=$?" >> gpurun_out/c_wide_test.log
tail -25 gpurun_out/c_wide_test.log
timeout 300 python tools/eig_perf.py > gpurun_out/c_eig_perf.log 2>&1
grep -E "symeig n=|^\{'n'|random|nvecs" gpurun_out/c_eig_perf.log | head -30
