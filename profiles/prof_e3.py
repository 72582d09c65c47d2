"""Pair marginals of ONE open train (axis e3 of SURVEY 8e): per-kernel CUDA-event times of twovar_marginals at
config-2, config-4 and config-3 size; k_twovar_open (a group of first sites per CTA) against the older kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
def mk(seed, L, d, q):
    rng = np.random.Generator(np.random.Philox(seed)); bonds=[1]+[d]*(L-1)+[1]
    return [rng.random((bonds[t], bonds[t+1], q)) for t in range(L)]
for (L,d,q) in [(128,64,2),(256,128,4),(256,256,2)]:
    A = T.TensorTrain(mk(2,L,d,q)); T.normalize_(A)
    T.twovar_marginals_block(A, 0, L-1)
    t0=time.perf_counter(); o=T.twovar_marginals_block(A, 0, L-1); w=time.perf_counter()-t0
    print(L,d,q,o.shape, "device ms %.3f" % A.last_timing()[0], "wall ms %.3f" % (w*1e3), "sums ok", np.allclose(o.sum(axis=2),1.0,atol=1e-9))
    A.profile(True); T.twovar_marginals_block(A, 0, L-1); rep=A.profile_report(); A.profile(False)
    for k,(n,ms) in sorted(rep.items(), key=lambda kv:-kv[1][1]): print("     %-22s %3d launches %8.3f ms" % (k,n,ms))
    del A
