#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -25 gpurun_out/i_pytest.log
python - <<'PY'
import sys, time, os
sys.path.insert(0,'scikit-tensor_b200'); sys.path.insert(0,'.')
import numpy as np, torch
from sktensor import cp_als, dtensor
X=np.random.RandomState(0).rand(10,11,8)
for env in ('0','1'):
    os.environ['SKT_NO_RUN_LOOP']=env
    ts=[]
    for _ in range(7):
        np.random.seed(42); t=time.perf_counter(); P,fit,itr,_=cp_als(dtensor(X),3,init='random'); ts.append(time.perf_counter()-t)
    print('cfg1 NO_RUN_LOOP=%s: %.3f ms (itr %d fit %.12f)'%(env, np.median(ts[2:])*1e3, itr, fit))
PY
