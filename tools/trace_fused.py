import os, sys
os.environ['SKT_FUSED_TRACE']='1'
sys.path.insert(0,'scikit-tensor_b200')
import numpy as np, torch
from sktensor import _device as dev, _kernels as K
for rows,R,N in ((512,32,3),(160,64,4),(512,16,3)):
    rng=np.random.RandomState(0)
    G=dev.to_device(np.stack([(lambda f: f.T.dot(f))(rng.rand(rows,R)) for _ in range(N)]))
    M=dev.to_device(rng.rand(rows,R)); U,lam,info=dev.empty((rows,R)),dev.empty((R,)),dev.zeros((1,),np.int32)
    ip,fit,nx2=dev.zeros((1,)),dev.zeros((2,)),dev.to_device(np.array([1e6]))
    print(rows,R,file=sys.stderr)
    for _ in range(3):
        K.cp_update_fused(M,G,0,False,U,lam,info,ip=ip,normX2=nx2,fit=fit)
