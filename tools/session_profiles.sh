#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_symeig.py -q) 2>&1 | tail -2
timeout 300 python tools/eig_perf.py 2>&1 | grep -E "symeig n=|last pass|its_per" | cut -c1-260 | head -14
timeout 120 python tools/fp64_peak.py > gpurun_out/l_fp64_peak.json 2>gpurun_out/l_fp64.err; cat gpurun_out/l_fp64_peak.json | head -8
timeout 120 python tools/hooi_one.py 20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l_launches_hooi.csv python tools/hooi_one.py 3 > gpurun_out/l_ncu_hooi.log 2>&1
tail -2 gpurun_out/l_ncu_hooi.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mttkrp_coo_kernel -s 3 -c 3 -o gpurun_out/l_sparse200M -f python tools/sparse_one.py 200000000 > gpurun_out/l_ncu_sparse.log 2>&1
tail -2 gpurun_out/l_ncu_sparse.log
ls -la gpurun_out/l_*
