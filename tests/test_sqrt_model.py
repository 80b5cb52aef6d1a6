"""Square-root step == g2o Schur step (the algebra the CUDA path relies on)."""
import numpy as np

from oracle import refba
from sqrt_model import sqrt_step, householder_qr_damped


def test_householder_q1_is_orthonormal_and_reproduces_A():
    rng = np.random.default_rng(0)
    for m, lam in [(6, 1e-3), (9, 10.0), (36, 1e4), (3, 1e-8)]:
        Jl = rng.normal(0, 30, (m, 3))
        R, Q1o = householder_qr_damped(Jl, lam)
        # obs rows of Q1 times R must give back the obs rows of A; the Gram of the full Q1 is I
        np.testing.assert_allclose(Q1o @ R, Jl, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(R.T @ R, Jl.T @ Jl + lam * np.eye(3), rtol=1e-12)
        Q1d = np.sqrt(lam) * np.linalg.inv(R)   # damping rows of Q1
        np.testing.assert_allclose(Q1o.T @ Q1o + Q1d.T @ Q1d, np.eye(3), atol=1e-13)


def test_sqrt_step_equals_schur_step(synth):
    for seed, stereo in [(0, True), (1, False)]:
        prob = synth.small_window(seed, n_free=5, n_fixed=2, n_points=80, mean_track=4.0, stereo=stereo)
        r = refba.RefBA(prob)
        lam = 12.5
        s = r.schur_solve(lam, huber=1)
        lin = r.linearize_all(1)
        pose_slot = -np.ones(prob.n_pose, int)
        pose_slot[s["slot_pose"]] = np.arange(s["Np"])
        out = sqrt_step(prob, lin, lam, pose_slot)
        Np = s["Np"]
        np.testing.assert_allclose(out["bs"].ravel(), s["bschur"], rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(out["dp"].ravel(), s["x"][:6 * Np], rtol=1e-7, atol=1e-11)
        np.testing.assert_allclose(out["dl"].ravel(), s["x"][6 * Np:], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(np.concatenate([out["bp"].ravel(), out["bl"].ravel()]), s["b"], rtol=1e-10, atol=1e-8)
        # implicit operator == explicit reduced matrix
        e = np.zeros((Np, 6)); e[1, 2] = 1.0
        np.testing.assert_allclose(out["matvec"](e).ravel(), s["S"][:, 6 + 2], rtol=1e-9, atol=1e-6)
