#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
SKT_SWEEP_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err
grep "sweep trace" gpurun_out/m_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/m_bench_n$N.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['als_iters_per_sec'])
print({k:v['ms_per_iteration'] for k,v in d['strong_scaling'].items()})
PY
