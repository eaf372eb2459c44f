"""FP64 peaks of the B200 this pool hands out, for profiles/fp64_peak.json: the DMMA issue-rate probe of the library
(skt_probe_fp64_peaks: register-only mma.sync.m8n8k4.f64 loop, and a DFMA loop) and cuBLAS DGEMM 8192^3 via torch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _kernels as K  # noqa: E402

dmma, dfma = K.probe_fp64_peaks()
n = 8192
a = torch.rand(n, n, dtype=torch.float64, device='cuda')
b = torch.rand(n, n, dtype=torch.float64, device='cuda')
for _ in range(2):
    torch.matmul(a, b)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
out = {'gpu': torch.cuda.get_device_name(0), 'dmma_probe_tflops': dmma, 'dfma_probe_tflops': dfma,
       'cublas_dgemm_8192_tflops_best': 2.0 * n ** 3 / (min(ts) * 1e-3) / 1e12,
       'cublas_dgemm_8192_ms': ts,
       'how': 'skt_probe_fp64_peaks (csrc/common.cu): register-only DMMA.8x8x4 / DFMA issue loops on all SMs, CUDA events; '
              'torch.matmul float64 8192^3, best of 5',
       'used_as': 'denominator of roofline.frac for the FP64-tensor-pipe-bound dense MTTKRP (MEASURED_PEAKS.json carries no '
                  'FP64 figure); bench.py re-measures the DMMA probe live on every run'}
print(json.dumps(out, indent=1))
