#!/bin/bash
# round-2 GPU session A: full test-suite (with the full-size parity tests), sparse variants, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q --durations=25) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --workload sparse --nnz 200000000 --steps 10 --variants > gpurun_out/a_sparse200M.json 2> gpurun_out/a_sparse200M.err
timeout 300 python bench.py --workload sparse --nnz 20000000 --steps 10 --variants > gpurun_out/a_sparse20M.json 2> gpurun_out/a_sparse20M.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -3 gpurun_out/a_pytest.log
