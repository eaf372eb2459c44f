#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_symeig.py tests/test_gpu_sparse_ops.py tests/test_gpu_widerank.py -q --durations=5) > gpurun_out/f_tests.log 2>&1
echo "rc=$?" >> gpurun_out/f_tests.log
tail -40 gpurun_out/f_tests.log
timeout 300 python tools/eig_perf.py > gpurun_out/f_eig_perf.log 2>&1
grep -E "symeig n=|^\{'n'|random|nvecs" gpurun_out/f_eig_perf.log | head -30
