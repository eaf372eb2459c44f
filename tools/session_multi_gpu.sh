#!/bin/bash
# multi-GPU session: NCCL tests + the default bench line at N ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
(time timeout 900 python -m pytest tests/test_gpu_multi.py -q -x) > gpurun_out/j_multi_n$N.log 2>&1
tail -6 gpurun_out/j_multi_n$N.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/j_bench_n$N.json 2> gpurun_out/j_bench_n$N.err
tail -4 gpurun_out/j_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/j_bench_n$N.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e it/s', d['e2e']['als_iters_per_sec'], d['e2e']['runs_s'])
print(d.get('strong_scaling')); print(d.get('extras_error'))
o=d.get('opt_in_int8_slices') or {}
print('opt-in int8 slices:', {k: o.get(k) for k in ('ms_per_step', 'als_iters_per_sec', 'speedup_vs_default', 'fit_delta_vs_default', 'e2e_als_iters_per_sec', 'error')})
PY
