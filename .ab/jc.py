import sys
sys.path[:0] = ["/root/repo", "/root/repo/tensortrains.jl_b200"]
import numpy as np, ttb200 as T
rng = np.random.Generator(np.random.Philox(3)); L,d,q=24,256,2
bonds=[1]+[d]*(L-1)+[1]
A=T.TensorTrain([rng.standard_normal((bonds[t],bonds[t+1],q))/16 for t in range(L)])
T.compress_(A, T.TruncThresh(1e-10)); print(A.last_timing())
