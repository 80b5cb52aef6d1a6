"""The kernels' __host__ __device__ arithmetic (csrc/sqrtba_math.cuh), compiled for the host, against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import refba

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mc():
    out = os.path.join(HERE, "_build", "libmathcheck.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "cpu_math_check.cpp")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src],
                   check=True)
    return C.CDLL(out)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("stereo", [True, False])
def test_residual_and_jacobians_match_oracle(mc, synth, stereo):
    prob = synth.small_window(4, n_points=120, stereo=stereo)
    lin = refba.RefBA(prob).linearize_all(0)
    e, Jp, Jl = np.zeros(3), np.zeros(18), np.zeros(9)
    dpos = C.c_int(0)
    for k in range(prob.n_obs):
        pose = np.ascontiguousarray(prob.pose_qt[prob.obs_pose[k]])
        X = np.ascontiguousarray(prob.point_xyz[prob.obs_point[k]])
        cam = np.ascontiguousarray(prob.cam[prob.obs_pose[k]])
        meas = np.ascontiguousarray(prob.obs_meas[k])
        mc.mc_obs(_dp(pose), _dp(X), _dp(cam), meas.ctypes.data_as(C.POINTER(C.c_float)), _dp(e), _dp(Jp), _dp(Jl),
                  C.byref(dpos))
        # float32 inverse depth: a last-bit difference in z may flip the rounding (1 float ulp of 1/z ~ 1e-5 px)
        assert np.abs(e - lin["err"][k]).max() <= (2e-5 if stereo else 1e-10)
        np.testing.assert_allclose(Jp.reshape(3, 6), lin["Jp"][k], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(Jl.reshape(3, 3), lin["Jl"][k], rtol=1e-12, atol=1e-10)
        assert dpos.value == 1


def test_pose_oplus_matches_oracle(mc):
    rng = np.random.default_rng(0)
    for scale in (0.3, 1e-3, 1e-7, 0.0):
        for _ in range(10):
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            if q[3] < 0:
                q = -q
            pose = np.concatenate([rng.normal(0, 30, 3), q])
            xi = rng.normal(0, 1, 6) * scale
            want = refba.pose_oplus(pose, xi)
            got = pose.copy()
            mc.mc_oplus(_dp(got), _dp(np.ascontiguousarray(xi)))
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_huber_matches_oracle(mc):
    d = float(np.float32(np.sqrt(7.815)))
    for c in (0.0, 1.0, d * d, d * d + 1e-9, 50.0, 1e6):
        r0, r1 = C.c_double(0), C.c_double(0)
        mc.mc_huber(C.c_double(c), C.c_double(d), C.byref(r0), C.byref(r1))
        want = np.zeros(3)
        refba.lib().refba_huber(d, c, _dp(want))
        # last-bit slack: the oracle build may contract 2*s*delta - dsqr into an FMA
        assert abs(r0.value - want[0]) <= 4e-16 * max(1.0, abs(want[0])) and r1.value == want[1]


def test_spd6_inverse(mc):
    rng = np.random.default_rng(1)
    for _ in range(10):
        B = rng.normal(size=(9, 6))
        A = np.ascontiguousarray(B.T @ B + 1e-3 * np.eye(6))
        Ai = np.zeros((6, 6))
        assert mc.mc_spd6_inverse(_dp(A), _dp(Ai)) == 1
        np.testing.assert_allclose(Ai @ A, np.eye(6), atol=1e-9)
    A = np.ascontiguousarray(-np.eye(6))
    assert mc.mc_spd6_inverse(_dp(A), _dp(np.zeros((6, 6)))) == 0


def test_sim3_arithmetic_matches_reference_binary(mc):
    """csrc/sqrtba_sim3.cuh (the arithmetic of the future essential-graph kernel) compiled for the host,
    against vectors recorded from the reference's own binary (oracle/pin_libg2o_graph.py: make_sim3)."""
    gold = np.load(os.path.join(HERE, "golden", "libg2o_vectors.npz"))
    for u, want in zip(gold["sim3_upd"], gold["sim3_exp"]):
        out = np.zeros(8)
        mc.mc_sim3_exp(_dp(np.ascontiguousarray(u)), _dp(out))
        np.testing.assert_allclose(out, want, rtol=0, atol=4e-15 * max(1.0, np.abs(want).max()))
    for s8, u, free, fix in zip(gold["sim3_exp"], gold["sim3_upd2"], gold["sim3_oplus_free"], gold["sim3_oplus_fix"]):
        for flag, want in ((0, free), (1, fix)):
            est = np.ascontiguousarray(s8).copy()
            mc.mc_sim3_oplus(_dp(est), _dp(np.ascontiguousarray(u)), flag)
            np.testing.assert_allclose(est, want, rtol=0, atol=4e-15 * max(1.0, np.abs(want).max()))
    for m, a, b, want in zip(gold["sim3_meas"], gold["sim3_v1"], gold["sim3_v2"], gold["sim3_err"]):
        out = np.zeros(7)
        mc.mc_sim3_edge_error(_dp(np.ascontiguousarray(m)), _dp(np.ascontiguousarray(a)), _dp(np.ascontiguousarray(b)), _dp(out))
        np.testing.assert_allclose(out, want, rtol=0, atol=1e-13 * max(1.0, np.abs(want).max()))
        back = np.zeros(8)                                   # exp(log(.)) closes on the same element
        mc.mc_sim3_exp(_dp(out), _dp(back))
        lg = np.zeros(7)
        mc.mc_sim3_log(_dp(back), _dp(lg))
        np.testing.assert_allclose(lg, out, rtol=0, atol=1e-12 * max(1.0, np.abs(out).max()))


@pytest.mark.parametrize("case", [0, 1])
def test_sim3_edge_jacobians_match_reference_binary(mc, case):
    """The numeric Jacobians EdgeSim3 inherits (central differences with delta = 1e-9 through oplusImpl), as the binary
    writes them into its JacobianWorkspace for every edge of the pose-graph fixtures, against sim3_edge_linearize.  The
    scheme amplifies rounding by 1/(2 delta) = 5e8, so two implementations agree to ~1e-5 (entries reach ~10), not to rounding."""
    gold = np.load(os.path.join(HERE, "golden", "libg2o_vectors.npz"))
    g = {k[len(f"pg{case}_"):]: gold[k] for k in gold.files if k.startswith(f"pg{case}_")}
    fix_scale = int(g["fix_scale"])
    seen_free = 0
    for (i, j), m, e0, Ji0, Jj0 in zip(g["edges"], g["meas"], g["err0"], g["Ji0"], g["Jj0"]):
        a, b = np.ascontiguousarray(g["vert0"][i]), np.ascontiguousarray(g["vert0"][j])
        err = np.zeros(7)
        mc.mc_sim3_edge_error(_dp(np.ascontiguousarray(m)), _dp(a), _dp(b), _dp(err))
        np.testing.assert_allclose(err, e0, rtol=0, atol=1e-13)
        Ji, Jj = np.zeros((7, 7)), np.zeros((7, 7))
        mc.mc_sim3_edge_linearize(_dp(np.ascontiguousarray(m)), _dp(a), _dp(b), int(g["fixed"][i]), int(g["fixed"][j]),
                                  fix_scale, _dp(Ji), _dp(Jj))
        np.testing.assert_allclose(Ji, Ji0, rtol=0, atol=2e-5)
        np.testing.assert_allclose(Jj, Jj0, rtol=0, atol=2e-5)
        if not g["fixed"][i]:
            seen_free += 1
            assert np.abs(Ji0).max() > 0.5
            assert (np.abs(Ji0[:, 6]).max() == 0.0) == bool(fix_scale)      # a fixed scale zeroes the last column
    assert seen_free > 10
