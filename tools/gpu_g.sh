#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
tail -5 gpurun_out/g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','als_iters_per_sec'): print(k, d[k])
print('e2e', d['e2e']['als_iters_per_sec']); print('cpu', d['cpu_baseline'])
for k in ('strong_scaling','cfg4_hooi_384^3_rank32','cfg1_latency','extras_error'):
    print(k, d.get(k))
rs=d.get('roofline_sparse')
if rs: print({m:(v['ms'],round(v['frac'],3)) for m,v in rs['modes'].items()})
PY
timeout 300 python tools/eig_perf.py 2>&1 | grep -E "symeig n=" | head -3
