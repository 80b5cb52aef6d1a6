"""NumPy model of the two-level preconditioner of global BA (csrc/sqrtba_chunkprec.cuh), checked against the oracle's
EXPLICIT reduced camera system on the CPU: the chunk blocks the kernels assemble from G = Jp^T Q1 are the principal
submatrices of S, the coarse matrix they assemble from block sums and cross-chunk pair sums is Z^T S Z, the branch-free
Gauss-Jordan sweep inverts, and PCG with the resulting operator reaches the oracle's damped step in fewer iterations
than with the 6x6 blocks.  Test infrastructure (the GPU tests compare the kernels themselves with the oracle)."""
import importlib

import numpy as np

from oracle import refba
from sqrt_model import householder_qr_damped

synth = importlib.import_module("sqrtlm-slam_b200.synth")
VSLOT = 20


def linearised(prob, lam):
    lin = refba.RefBA(prob).linearize_all(0)
    sw = np.sqrt(lin["w"])[:, None]
    Jp, Jl = lin["Jp"] * sw[:, :, None], lin["Jl"] * sw[:, :, None]
    ptr = prob.lm_ptr()
    G = np.zeros((prob.n_obs, 6, 3))
    for l in range(prob.n_point):
        a, b = ptr[l], ptr[l + 1]
        _, Q = householder_qr_damped(Jl[a:b].reshape(-1, 3), lam)
        Q1 = Q.reshape(-1, 3, 3)
        G[a:b] = np.einsum("kri,krj->kij", Jp[a:b], Q1)      # Jp^T Q1, 6x3 per observation
    return Jp, G, ptr


def sweep_inverse(A):
    """k_chunk_factor's sweep: Jacobi scaling, then one rank-one update per pivot with the special cases folded in."""
    d = 1.0 / np.sqrt(np.diag(A))
    T = A * d[:, None] * d[None, :]
    n = len(T)
    for k in range(n):
        p = T[k, k]
        assert p > 0
        c = T[:, k].copy()
        r = T[k, :].copy()
        c[k] = p - 1.0
        r[k] = p + 1.0
        T = T - np.outer(c, r / p)
    return T * d[:, None] * d[None, :]


def test_chunk_blocks_coarse_matrix_and_sweep_match_the_explicit_reduced_system():
    prob = synth.make_problem(5, 58, 1, 1400, 8.0, stereo=True, loop=True, cand_halfwidth=12, name="prec-model")
    lam = 3.7
    s = refba.RefBA(prob).schur_solve(lam, huber=0)
    S, Np = s["S"], s["Np"]
    free = np.nonzero(prob.pose_fixed == 0)[0]
    slot = -np.ones(prob.n_pose, int)
    slot[free] = np.arange(Np)
    Jp, G, ptr = linearised(prob, lam)
    so = slot[prob.obs_pose]
    nch = (Np + VSLOT - 1) // VSLOT
    # ---- what the kernels accumulate
    D = np.tile(lam * np.eye(6), (Np, 1, 1))                 # landmark QR: diagonal 6x6 blocks
    for k in np.nonzero(so >= 0)[0]:
        D[so[k]] += Jp[k].T @ Jp[k] - G[k] @ G[k].T
    M = np.zeros((nch, VSLOT * 6, VSLOT * 6))                # k_chunk_blocks: one-sided pair blocks inside a chunk
    Cacc = np.zeros((nch, nch, 6, 6))                        # ... and cross-chunk pair sums
    for l in range(prob.n_point):
        ks = [k for k in range(ptr[l], ptr[l + 1]) if so[k] >= 0]
        for ii, i in enumerate(ks):
            for j in ks[ii + 1:]:
                ci, cj = so[i] // VSLOT, so[j] // VSLOT
                blk = -G[i] @ G[j].T
                if ci == cj:
                    ri, rj = (so[i] - ci * VSLOT) * 6, (so[j] - cj * VSLOT) * 6
                    M[ci, ri:ri + 6, rj:rj + 6] += blk
                else:
                    Cacc[ci, cj] += blk
    # ---- k_chunk_factor's assembly = principal submatrix of the oracle's S
    blocks = []
    for c in range(nch):
        n = (min(VSLOT, Np - c * VSLOT)) * 6
        A = M[c, :n, :n] + M[c, :n, :n].T
        for b in range(n // 6):
            A[b * 6:b * 6 + 6, b * 6:b * 6 + 6] = D[c * VSLOT + b]
        want = S[c * VSLOT * 6:c * VSLOT * 6 + n, c * VSLOT * 6:c * VSLOT * 6 + n]
        np.testing.assert_allclose(A, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
        inv = sweep_inverse(A)
        np.testing.assert_allclose(inv @ A, np.eye(n), atol=1e-9)
        blocks.append((A, inv))
    # ---- coarse matrix: diagonal blocks = 6x6 sums of the chunk matrices, off-diagonal = cross sums (+ transposes)
    Z = np.zeros((6 * Np, 6 * nch))
    for i in range(Np):
        Z[6 * i:6 * i + 6, 6 * (i // VSLOT):6 * (i // VSLOT) + 6] = np.eye(6)
    Ac = np.zeros((6 * nch, 6 * nch))
    for c in range(nch):
        A = blocks[c][0]
        n = A.shape[0]
        Ac[6 * c:6 * c + 6, 6 * c:6 * c + 6] = A.reshape(n // 6, 6, n // 6, 6).sum(axis=(0, 2))
        for c2 in range(nch):
            if c2 != c:
                Ac[6 * c:6 * c + 6, 6 * c2:6 * c2 + 6] = Cacc[c, c2] + Cacc[c2, c].T
    want = Z.T @ S @ Z
    np.testing.assert_allclose(Ac, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())

    # ---- PCG on the oracle's system with the operators the kernels apply (chunk inverses rounded to FP32)
    Aci = np.linalg.inv(Ac)

    def pcg(Minv, tol=1e-9):
        b = s["bschur"]
        x = np.zeros_like(b)
        r = b.copy()
        z = Minv(r)
        p = z.copy()
        rz = rz0 = r @ z
        it = 0
        while it < 2000 and rz > tol * tol * rz0:
            q = S @ p
            a = rz / (p @ q)
            x += a * p
            r -= a * q
            z = Minv(r)
            rzn = r @ z
            p = z + (rzn / rz) * p
            rz = rzn
            it += 1
        return x, it

    def m_jacobi(v):
        return np.einsum("nij,nj->ni", np.linalg.inv(D), v.reshape(Np, 6)).ravel()

    def m_chunk(v):
        out = np.empty_like(v)
        for c, (_, inv) in enumerate(blocks):
            a = c * VSLOT * 6
            out[a:a + len(inv)] = inv.astype(np.float32).astype(np.float64) @ v[a:a + len(inv)]
        return out

    def m_two_level(v):
        return m_chunk(v) + Z @ (Aci @ (Z.T @ v))

    xp = s["x"][:6 * Np]
    x6, it6 = pcg(m_jacobi)
    xc, itc = pcg(m_chunk)
    x2, it2 = pcg(m_two_level)
    for x in (x6, xc, x2):
        assert np.abs(x - xp).max() <= 1e-6 * np.abs(xp).max()
    assert it2 <= itc < it6, (it6, itc, it2)
    print(f"PCG iterations on a {Np}-keyframe loop: 6x6 blocks {it6}, chunk blocks {itc}, chunk + coarse {it2}")


def test_sweep_step_is_gauss_jordan_without_special_cases():
    rng = np.random.default_rng(0)
    for n in (6, 30, 120):
        B = rng.normal(size=(n, n + 5))
        ds = 10.0 ** rng.uniform(-3, 3, n)                       # badly scaled SPD, like rotations against translations
        A = (B @ B.T / n + 1e-2 * np.eye(n)) * np.outer(ds, ds)
        inv = sweep_inverse(A)
        np.testing.assert_allclose(inv @ A, np.eye(n), atol=1e-7)
        np.testing.assert_allclose(inv, inv.T, rtol=1e-8, atol=1e-12 * np.abs(inv).max())
