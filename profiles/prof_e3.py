import os, sys, time
sys.path[:0] = ["/root/repo", "/root/repo/tensortrains.jl_b200"]
import numpy as np
import ttb200 as T
def mk(seed, L, d, q):
    rng = np.random.Generator(np.random.Philox(seed)); bonds=[1]+[d]*(L-1)+[1]
    return [rng.random((bonds[t], bonds[t+1], q)) for t in range(L)]
for (L,d,q) in [(128,64,2),(256,128,4),(256,256,2)]:
    A = T.TensorTrain(mk(2,L,d,q)); T.normalize_(A)
    for lo,hi in [(0,L-1),(0,(L-1)//4)]:
        T.twovar_marginals_block(A, lo, hi)
        t0=time.perf_counter(); o=T.twovar_marginals_block(A, lo, hi); w=time.perf_counter()-t0
        print(L,d,q,(lo,hi),o.shape, "device ms", A.last_timing(), "wall ms", w*1e3, "sum ok", np.allclose(o.sum(axis=2),1.0,atol=1e-9))
    del A
