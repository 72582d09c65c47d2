import numpy as np, sys, math
sys.path.insert(0, "/root/repo")
rng = np.random.Generator(np.random.Philox(3))

def rr_pairs(nvp, rd):
    out = []
    for q in range(nvp // 2):
        if q == 0: i, j = nvp - 1, rd
        else:
            i = rd + q;  i -= (nvp - 1) if i >= nvp - 1 else 0
            j = rd - q;  j += (nvp - 1) if j < 0 else 0
        if i > j: i, j = j, i
        out.append((i, j))
    return out

def jacobi_sweeps(B, sort=False, maxsw=60, skip_tiny=None):
    B = B.copy(); nv, ln = B.shape
    if sort:
        o = np.argsort(-np.sum(B * B, axis=1), kind="stable"); B = B[o]
    nvp = (nv + 1) & ~1
    tol2 = max(ln, 2) * (2.220446049250313e-16) ** 2
    nrot_hist = []
    for sw in range(maxsw):
        sq = np.sum(B * B, axis=1)
        nrot = 0
        for rd in range(nvp - 1):
            pr = [(i, j) for i, j in rr_pairs(nvp, rd) if j < nv]
            I = np.array([p[0] for p in pr]); J = np.array([p[1] for p in pr])
            al, be = sq[I], sq[J]
            ga = np.sum(B[I] * B[J], axis=1)
            m = (al > 0) & (be > 0) & (ga * ga > tol2 * al * be)
            if skip_tiny is not None:
                mx = sq.max()
                m &= ~((al < skip_tiny * mx) & (be < skip_tiny * mx))
            if not m.any(): continue
            I, J, al, be, ga = I[m], J[m], al[m], be[m], ga[m]
            da = be - al; b2 = 2 * ga
            t = b2 / (da + np.copysign(np.sqrt(da * da + b2 * b2), da))
            cs = 1 / np.sqrt(1 + t * t); sn = cs * t
            x, y = B[I].copy(), B[J].copy()
            B[I] = cs[:, None] * x - sn[:, None] * y
            B[J] = sn[:, None] * x + cs[:, None] * y
            sq[I] = al - t * ga; sq[J] = be + t * ga
            nrot += len(I)
        nrot_hist.append(nrot)
        if nrot == 0: break
    return len(nrot_hist), nrot_hist, np.sort(np.sqrt(np.sum(B * B, axis=1)))[::-1]

def left_sweep_mats(L, d, q, rule, fill="rand"):
    """eps=0 right sweep (QR) then truncating left sweep; yield the matrices handed to the SVD."""
    bonds = [1] + [d] * (L - 1) + [1]
    f = (lambda s: rng.random(s)) if fill == "rand" else (lambda s: rng.standard_normal(s) / math.sqrt(d))
    A = [f((bonds[t], bonds[t + 1], q)) for t in range(L)]
    for t in range(L - 1, 0, -1):
        dl, dr, _ = A[t].shape
        M = A[t].reshape(dl, dr * q, order="F")           # dl x (dr q)
        Q_, R = np.linalg.qr(M.T)                          # M^T = Q R ; M = R^T Q^T
        k = Q_.shape[1]
        A[t] = Q_.T.reshape(k, dr, q, order="F")
        S = R.T                                            # dl x k
        nxt = np.einsum("abx,bc->acx", A[t - 1], S)
        A[t - 1] = nxt / np.abs(nxt).max()
    mats = []
    for t in range(L - 1):
        dl, dr, _ = A[t].shape
        X = np.concatenate([A[t][:, :, x] for x in range(q)], axis=0)   # (dl q) x dr, row = a + dl*x
        mats.append(X.copy())
        U, s, Vt = np.linalg.svd(X, full_matrices=False)
        mp = rule(s)
        A[t] = np.stack([U[x * dl:(x + 1) * dl, :mp] for x in range(q)], axis=2)
        S = s[:mp, None] * Vt[:mp]
        nxt = np.einsum("ab,bcx->acx", S, A[t + 1])
        A[t + 1] = nxt / np.abs(nxt).max()
    return mats

def thresh(eps):
    def r(s):
        tot = math.sqrt(np.sum(s * s)) * eps; thr = tot * tot
        acc = s[-1] ** 2; mp = 1
        for k in range(len(s) - 1, 0, -1):
            acc += s[k - 1] ** 2
            if acc >= thr: mp = k + 1; break
        return mp
    return r

def study(name, mats, idx):
    for t in idx:
        X = mats[t]; p, n = X.shape
        if p >= n:
            R = np.linalg.qr(X, mode="r"); Bm = R; kind = "rows of R"
        else:
            Bm = X; kind = "rows of X"
        a = jacobi_sweeps(Bm); b = jacobi_sweeps(Bm, sort=True); c = jacobi_sweeps(Bm, sort=True, skip_tiny=1e-28)
        e = jacobi_sweeps(Bm.T.copy() if p >= n else Bm, sort=True)
        print(f"{name} t={t} {p}x{n} {kind}: plain {a[0]} {a[1][:12]} | sorted {b[0]} {b[1][:12]} | sorted+skip {c[0]} | transposed-sorted {e[0]}")

which = sys.argv[1]
if which == "c3":
    mats = left_sweep_mats(12, 128, 2, thresh(1e-10)); study("c3-like d=128", mats, [2, 5, 8])
elif which == "c2":
    mats = left_sweep_mats(12, 64, 2, lambda s: min(len(s), 32)); study("c2", mats, [2, 5, 8])
elif which == "randn":
    mats = left_sweep_mats(8, 64, 2, thresh(1e-10), "randn"); study("randn", mats, [2, 4])
