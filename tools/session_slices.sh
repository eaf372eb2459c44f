#!/bin/bash
# sliced INT8 contraction: tests, timings, ncu launch list and one full capture of slice_gemm_kernel (both majors)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_slices.py -q -m gpu) 2>&1 | tail -2
timeout 120 python tools/slices_perf.py 2>&1 | tee gpurun_out/m_slices_perf.log
timeout 120 python tools/slices_one.py > gpurun_out/m_one.log 2>&1 || { tail -5 gpurun_out/m_one.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/m_launches_slices.csv python tools/slices_one.py > gpurun_out/m_ncu_l.log 2>&1
tail -2 gpurun_out/m_ncu_l.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:slice_gemm_kernel -s 2 -c 2 -o gpurun_out/m_slices -f python tools/slices_one.py > gpurun_out/m_ncu_f.log 2>&1
tail -2 gpurun_out/m_ncu_f.log
ls -la gpurun_out/m_*
