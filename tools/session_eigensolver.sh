#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_symeig.py -q) > gpurun_out/h_tests.log 2>&1
tail -3 gpurun_out/h_tests.log
timeout 300 python tools/eig_perf.py 2>&1 | grep -E "symeig n=|random 20|nvecs 20" | head -8
