#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_sparse_ops.py tests/test_gpu_fuzz.py -q -k "not nvecs and not default_init") > gpurun_out/e_sparse_ops.log 2>&1
echo "rc=$?" >> gpurun_out/e_sparse_ops.log
tail -30 gpurun_out/e_sparse_ops.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:symeig -c 1 -o gpurun_out/e_symeig384 -f python tools/eig_one.py 384 1024 32 > gpurun_out/e_ncu.log 2>&1
tail -3 gpurun_out/e_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -2
