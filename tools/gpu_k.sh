#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python tools/eig_perf.py 2>&1 | grep -E "symeig n=" | head -8
SKT_EIG_NO_CHOLQR=1 timeout 300 python tools/eig_perf.py 2>&1 | grep -E "symeig n=" | head -3
