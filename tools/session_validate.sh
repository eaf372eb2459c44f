#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -6 gpurun_out/v_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -2 gpurun_out/v_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/v_ncu_bench.log 2>&1
tail -1 gpurun_out/v_ncu_bench.log | cut -c1-200
python -c "
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['als_iters_per_sec'], d['roofline']['frac'], d['cpu_baseline']['kind'], d.get('extras_error'))
print({k:round(v['ms_per_iteration'],3) for k,v in d['strong_scaling'].items()}, d['cfg4_hooi_384^3_rank32']['nvecs'], d['cfg1_latency']['gpu_whole_call_ms'])
"
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python tools/eig_perf.py 2>&1 | grep -E "symeig n=" | cut -c1-170
