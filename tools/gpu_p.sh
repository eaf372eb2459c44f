#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
(timeout 900 python -m pytest tests/test_gpu_mttkrp.py tests/test_gpu_sparse_ops.py tests/test_gpu_rescal.py tests/test_gpu_fullsize.py tests/test_gpu_cpals.py tests/test_gpu_fuzz.py -x -q -k "sparse or coo or cfg3 or accum or compact or rescal or nvecs or empty or kat3 or sp_") 2>&1 | tail -4
timeout 300 python bench.py --workload sparse --nnz 200000000 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('build s', d['build_seconds_device_arrays'], {m:round(v['ms'],3) for m,v in d['modes'].items()})"
