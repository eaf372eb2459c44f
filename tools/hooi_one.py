"""tucker_hooi(384^3, [32,32,32]) for a few iterations -- the ncu launch-list target of BASELINE configs[3]."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import dtensor, tucker_hooi  # noqa: E402

X = dtensor(np.random.RandomState(0).rand(384, 384, 384))
np.random.seed(0)
tic = time.perf_counter()
core, U = tucker_hooi(X, [32, 32, 32], init='nvecs', maxIter=int(sys.argv[1]) if len(sys.argv) > 1 else 3, conv=0)
torch.cuda.synchronize()
print('seconds', time.perf_counter() - tic, 'core norm', float(np.linalg.norm(core)))
