"""Config 5 at batch B (default 4096): per-kernel CUDA-event times of normalize!, marginals,
twovar_marginals and compress!(TruncThresh(1e-6)) on a batch of periodic trains (64 sites, bond 32, q=2).
Packed upload (no per-train Python objects).  Timing only: parity lives in tests/."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L, d, q = 64, 32, 2
site = d * d * q
rng = np.random.Generator(np.random.Philox(5))
packed = rng.random(B * L * site)
A = T.PeriodicTensorTrain.empty([d] * (L + 1), (q,), batch=B)
A.upload_all(packed)
nbytes = packed.nbytes
print(f"C5 B={B}: slab {nbytes/2**30:.2f} GiB")


def run(name, f, alg_bytes=None):
    W = A.copy()
    f(W)                      # warm-up (workspace allocation)
    W = A.copy()
    W.profile(True)
    r = f(W)
    rep = W.profile_report()
    W.profile(False)
    W2 = A.copy()
    f(W2) if name != "compress!" else None
    ms, nl = (W2 if name != "compress!" else W).last_timing()
    print(f"{name}: device {ms:.2f} ms (unprofiled), launches {nl}")
    for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        extra = ""
        if alg_bytes and k in alg_bytes:
            extra = f"  {alg_bytes[k]/(t*1e-3)/1e9:.0f} GB/s algorithmic"
        print(f"    {k:24s} {n:5d} launches {t:9.3f} ms{extra}")
    return r, W


lz, Wn = run("normalize!", lambda W: T.normalize_(W),
             {"k_accumulate_L": nbytes, "k_env_batch": nbytes, "k_scale_trains": 2 * nbytes})
def raw_marg(W):
    out = np.zeros((B, L, q))
    T._ck(T.lib.ttb_marginals(W._hd, T._dp(out)))
    return out
_, _ = run("marginals", raw_marg)
def raw_twovar(W):
    import ctypes as C
    n = C.c_int64()
    T._ck(T.lib.ttb_twovar_count(W._hd, L - 1, C.byref(n)))
    out = np.zeros((B, n.value, q * q))
    T._ck(T.lib.ttb_twovar_marginals(W._hd, L - 1, T._dp(out)))
    return out
_, _ = run("twovar_marginals", raw_twovar)
_, _ = run("compress!", lambda W: T.compress_(W, T.TruncThresh(1e-6)))
