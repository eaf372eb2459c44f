#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
(timeout 600 python -m pytest tests/test_gpu_cpals.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_gpu_sparse_ops.py -q -k "hooi or tucker or ttm") 2>&1 | tail -3
timeout 300 python tools/eig_perf.py 2>&1 | grep -E "its_per" 
