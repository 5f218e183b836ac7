"""Numerical model (numpy) of the blocked Jacobi eigensolver of deepdow_b200/csrc/eig_block.cu.

It mirrors the kernel's arithmetic closely enough to predict accuracy before spending GPU time: TF32 hi/lo split of both
operands, tcgen05 accumulation modelled as round-toward-zero after every k-step of 8 (the model reproduces the error the
probe measured on a B200: 1.4e-5 predicted vs 1.2e-5 measured at K=250), separate accumulators for the cross terms,
rotation updates in "delta" form X + X (U - I).
"""
import sys
import numpy as np


def trunc32(x):
    y = x.astype(np.float32)
    big = np.abs(y.astype(np.float64)) > np.abs(x)
    y[big] = np.nextafter(y[big], np.float32(0))
    return y


def tf32_rna(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def mma3(A, B, nacc=3, exact=False):
    """C = A @ B.T as the tile GEMM computes it. A (M,K), B (N,K) float32."""
    if exact:
        return (A.astype(np.float64) @ B.astype(np.float64).T).astype(np.float32)
    Ah = tf32_rna(A); Al = tf32_rna(A - Ah)
    Bh = tf32_rna(B); Bl = tf32_rna(B - Bh)
    M, K = A.shape
    N = B.shape[0]
    accs = [np.zeros((M, N), np.float32) for _ in range(nacc)]
    cross = np.zeros((M, N), np.float32)
    f8 = np.float64
    for i, k0 in enumerate(range(0, K, 8)):
        sl = slice(k0, k0 + 8)
        cross = trunc32(cross.astype(f8) + Al[:, sl].astype(f8) @ Bh[:, sl].astype(f8).T)
        cross = trunc32(cross.astype(f8) + Ah[:, sl].astype(f8) @ Bl[:, sl].astype(f8).T)
        t = i % nacc
        accs[t] = trunc32(accs[t].astype(f8) + Ah[:, sl].astype(f8) @ Bh[:, sl].astype(f8).T)
    out = cross
    for a in accs:
        out = out + a
    return out


def local_sweeps(P, sweeps):
    """Parallel-order cyclic two-sided Jacobi on the symmetric float32 matrix P; returns U (columns rotate) and #rotations."""
    n = P.shape[0]
    S = P.astype(np.float32).copy()
    U = np.eye(n, dtype=np.float32)
    eps = np.float32(1.1920929e-7)
    nrot = 0
    idx = np.arange(n)
    for _ in range(sweeps):
        for r in range(n - 1):
            k = np.arange(n // 2)
            a = np.where(k == 0, r, (r + k) % (n - 1))
            b = np.where(k == 0, n - 1, (r - k + (n - 1)) % (n - 1))
            p = np.minimum(a, b); q = np.maximum(a, b)
            spq = S[p, q]; spp = S[p, p]; sqq = S[q, q]
            act = (np.abs(spq) > eps * np.sqrt(np.abs(spp * sqq))) & (np.abs(spq) > 1e-30)
            with np.errstate(divide="ignore", invalid="ignore"):
                tau = (sqq - spp) / (2 * spq)
                t = np.sign(tau + (tau == 0)) / (np.abs(tau) + np.sqrt(1 + tau * tau))
            t = np.where(act, t, 0).astype(np.float32)
            c = (1 / np.sqrt(1 + t * t)).astype(np.float32)
            s = (t * c).astype(np.float32)
            nrot += int((np.abs(spq) > 8 * eps * np.sqrt(np.abs(spp * sqq))).sum())
            tt = (s / (1 + c)).astype(np.float32)
            # Rutishauser's form a' = a - s (b + tau a), b' = b + s (a - tau b): keeps the 1 - c term in float32
            Sp = S[:, p].copy(); Sq = S[:, q].copy()
            S[:, p] = Sp - s * (Sq + tt * Sp); S[:, q] = Sq + s * (Sp - tt * Sq)
            Sp = S[p, :].copy(); Sq = S[q, :].copy()
            S[p, :] = Sp - s[:, None] * (Sq + tt[:, None] * Sp); S[q, :] = Sq + s[:, None] * (Sp - tt[:, None] * Sq)
            Up = U[:, p].copy(); Uq = U[:, q].copy()
            U[:, p] = Up - s * (Uq + tt * Up); U[:, q] = Uq + s * (Up - tt * Uq)
    return U, nrot


def block_jacobi(C, b=64, inner=1, max_sweeps=16, exact=False, delta=True, verbose=False):
    N = C.shape[0]
    nb = -(-N // b)
    nb += nb & 1
    NP = nb * b
    S = np.zeros((NP, NP), np.float32); S[:N, :N] = C.astype(np.float32)
    VT = np.eye(NP, dtype=np.float32)
    I2 = np.eye(2 * b, dtype=np.float32)
    for sweep in range(max_sweeps):
        total_rot = 0
        for r in range(nb - 1):
            pairs = []
            for k in range(nb // 2):
                a_, b_ = (r, nb - 1) if k == 0 else ((r + k) % (nb - 1), (r - k + nb - 1) % (nb - 1))
                pairs.append((min(a_, b_), max(a_, b_)))
            Us = []
            for (bi, bj) in pairs:
                ij = np.r_[bi * b:(bi + 1) * b, bj * b:(bj + 1) * b]
                U, nrot = local_sweeps(S[np.ix_(ij, ij)], inner)
                total_rot += nrot
                Us.append((ij, U))
            W = S.copy()
            for ij, U in Us:
                E = (U - I2) if delta else U
                # W[IJ, :] = U' S[IJ, :]  ==  tile GEMM  C(m=col, n) = sum_k S[IJ_k, m] * E[k, n], stored transposed
                upd = mma3(S[ij, :].T.copy(), E.T.copy(), exact=exact).T
                W[ij, :] = (S[ij, :] + upd) if delta else upd
                updv = mma3(VT[ij, :].T.copy(), E.T.copy(), exact=exact).T
                VT[ij, :] = (VT[ij, :] + updv) if delta else updv
            S2 = W.copy()
            for ij, U in Us:
                E = (U - I2) if delta else U
                # S_new[IJ, :] = U' W'[IJ, :] == C(m, n) = sum_k W[m, IJ_k] E[k, n], stored transposed
                upd = mma3(W[:, ij].copy(), E.T.copy(), exact=exact).T
                S2[ij, :] = (W[:, ij].T + upd) if delta else upd
            S = S2
        off = S - np.diag(np.diag(S))
        if verbose:
            print(f"  sweep {sweep}: rotations {total_rot}, off-norm {np.abs(off).max():.3e}, sym {np.abs(S - S.T).max():.3e}")
        if total_rot == 0:
            break
    return np.diag(S)[:N].copy(), VT[:N, :N].copy(), sweep + 1


def report(N, L, seed, **kw):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((L, N)) * 0.01
    Xc = X - X.mean(0)
    Sm = Xc.T @ Xc / (L - 1)
    C = 0.5 * Sm + 0.5 * np.diag(np.diag(Sm))
    lam, VT, sweeps = block_jacobi(C, verbose=True, **kw)
    V = VT.T.astype(np.float64)
    lam64, V64 = np.linalg.eigh(C)
    R64 = (V64 * np.sqrt(np.abs(lam64))) @ V64.T
    R = (V * np.sqrt(np.abs(lam.astype(np.float64)))) @ V.T
    ray = np.einsum("ik,ij,jk->k", V, C, V) / np.einsum("ik,ik->k", V, V)
    R2 = (V * np.sqrt(np.abs(ray))) @ V.T
    print(f"N={N} L={L} {kw}: sweeps {sweeps}  |V'V-I| {np.abs(V.T @ V - np.eye(N)).max():.2e}  "
          f"eig rel {np.abs(np.sort(lam) - lam64).max() / lam64.max():.2e}  rayleigh rel {np.abs(np.sort(ray) - lam64).max() / lam64.max():.2e}  "
          f"sqrt rel {np.abs(R - R64).max() / np.abs(R64).max():.2e}  sqrt(rayleigh) rel {np.abs(R2 - R64).max() / np.abs(R64).max():.2e}")


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    report(N, L, 0, exact=True)
    report(N, L, 0, exact=False, delta=True)
    report(N, L, 0, exact=False, delta=False)
