"""The oracle against the reference's OWN compiled code.

`oracle/pin_libg2o.py` calls the per-edge functions that the reference's prebuilt `Thirdparty/g2o/lib/libg2o.so` exports
(SE3Quat::exp, project2d, mono / stereo cam_project incl. the float32 inverse-depth quirk, RobustKernelHuber::robustify)
and stores inputs + outputs in `tests/golden/libg2o_vectors.npz` (committed; regenerate with `python oracle/pin_libg2o.py`
where /root/reference exists).  Here the oracle's restatements of those functions must reproduce the binary's outputs --
to a few ulps, since the binary and the oracle are compiled with different FMA contraction."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pin_libg2o, refba

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libg2o_vectors.npz")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def gold():
    assert os.path.exists(GOLD), "tests/golden/libg2o_vectors.npz is a committed fixture"
    return np.load(GOLD)


def test_se3_exp_matches_reference_binary(gold):
    L = refba.lib()
    for u, want in zip(gold["upd"], gold["exp"]):
        out = np.zeros(7)
        L.refba_se3_exp(_dp(np.ascontiguousarray(u)), _dp(out))
        np.testing.assert_allclose(out, want, rtol=0, atol=4e-15 * max(1.0, np.abs(want).max()))
    # both branches of se3quat.h:237-254 are in the fixture
    th = np.linalg.norm(gold["upd"][:, :3], axis=1)
    assert (th < 1e-5).sum() >= 10 and (th > 1e-2).sum() >= 100


def test_projection_matches_reference_binary(gold):
    L = refba.lib()
    cam = gold["cam"]
    for x, pm, ps, p2 in zip(gold["xyz"], gold["cam_mono"], gold["cam_stereo"], gold["project2d"]):
        x = np.ascontiguousarray(x)
        o2, o3 = np.zeros(2), np.zeros(3)
        L.refba_cam_project_mono(_dp(x), cam[0], cam[1], cam[2], cam[3], _dp(o2))
        np.testing.assert_allclose(o2, pm, rtol=4e-16, atol=0)
        np.testing.assert_allclose((o2 - cam[2:4]) / cam[:2], p2, rtol=0, atol=1e-13)
        L.refba_cam_project_stereo(_dp(x), cam[0], cam[1], cam[2], cam[3], C.c_float(cam[4]), _dp(o3))
        np.testing.assert_allclose(o3, ps, rtol=4e-16, atol=0)
    # the quirk is real: the binary's stereo u differs from the double-precision u by float rounding of 1/z
    du = np.abs(gold["cam_stereo"][:, 0] - gold["cam_mono"][:, 0])
    assert du.max() > 1e-7 and du.max() < 1e-3


def test_huber_matches_reference_binary(gold):
    L = refba.lib()
    delta = float(gold["delta"])
    for e2, want in zip(gold["e2"], gold["huber"]):
        rho = np.zeros(3)
        L.refba_huber(delta, float(e2), _dp(rho))
        np.testing.assert_allclose(rho, want, rtol=4e-16, atol=0)
    assert (gold["huber"][:, 1] == 1.0).sum() > 50 and (gold["huber"][:, 1] < 1.0).sum() > 50


@pytest.mark.skipif(not pin_libg2o.available(), reason="the reference checkout is not present on this machine")
def test_committed_fixture_is_what_the_binary_computes(tmp_path, gold):
    # in a clean interpreter: the prebuilt binary is not loaded into the test process
    out = str(tmp_path / "g.npz")
    odir = os.path.dirname(os.path.abspath(pin_libg2o.__file__))
    for script in ("pin_libg2o.py", "pin_libg2o_edges.py", "pin_libg2o_graph.py"):   # each appends its arrays
        subprocess.run([sys.executable, os.path.join(odir, script), out], check=True, capture_output=True)
    fresh = np.load(out)
    for k in gold.files:
        assert np.array_equal(fresh[k], gold[k]), k


def test_pose_oplus_chain_matches_reference_binary(gold):
    """VertexSE3Expmap::oplusImpl on a real vertex object of the binary: three updates from the origin
    (left-multiplication, quaternion product, normalizeRotation) -- the oracle's refba_pose_oplus and the kernels'
    pose_oplus (compiled for the host, tests/cpu_math_check.cpp) must land on the same pose."""
    L = refba.lib()
    for k, want in enumerate(gold["oplus"]):
        pose = np.array([0, 0, 0, 0, 0, 0, 1.0])
        for u in gold["upd"][3 * k:3 * k + 3]:
            out = np.zeros(7)
            L.refba_pose_oplus(_dp(pose), _dp(np.ascontiguousarray(u)), _dp(out))
            pose = out
        np.testing.assert_allclose(pose, want, rtol=0, atol=1e-14 * max(1.0, np.abs(want).max()))


@pytest.mark.parametrize("tag,stereo", [("mono", False), ("stereo", True)])
def test_edge_error_and_jacobians_match_reference_binary(gold, synth, tag, stereo):
    """REAL EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ objects of the binary (oracle/pin_libg2o_edges.py): computeError()
    and linearizeOplus() on random poses, points and measurements -- the oracle's residual and both Jacobians agree to
    rounding (the finite-difference check of test_oracle.py can only see the stereo Jacobian to 2e-2 because of the
    float32 inverse depth; this one is exact)."""
    d = 3 if stereo else 2
    cam = gold["edge_cam"]
    for k in range(len(gold[f"edge_{tag}_X"])):
        m = np.zeros((1, 4), np.float32)
        m[0, :3] = gold[f"edge_{tag}_meas"][k]
        m[0, 3] = 1.0
        if not stereo:
            m[0, 2] = -1.0
        prob = synth.Problem(gold[f"edge_{tag}_pose"][k][None].copy(), np.zeros(1, np.uint8), cam[None].copy(),
                             gold[f"edge_{tag}_X"][k][None].copy(), np.zeros(1, np.int32), np.zeros(1, np.int32), m)
        lin = refba.RefBA(prob).linearize_all(0)
        e, Jp, Jl = lin["err"][0][:d], lin["Jp"][0][:d], lin["Jl"][0][:d]
        np.testing.assert_allclose(e, gold[f"edge_{tag}_err"][k], rtol=0, atol=1e-14 * max(1.0, np.abs(e).max()))
        np.testing.assert_allclose(Jp, gold[f"edge_{tag}_Jp"][k], rtol=0, atol=4e-15 * np.abs(Jp).max())
        np.testing.assert_allclose(Jl, gold[f"edge_{tag}_Jl"][k], rtol=0, atol=4e-15 * np.abs(Jl).max())


def test_graph_semantics_match_reference_binary(gold, synth):
    """A REAL g2o::SparseOptimizer of the binary (oracle/pin_libg2o_graph.py) on a small local-BA-shaped graph, two
    phases like the two passes of local BA: (A) all edges at level 0 with Huber kernels, (B) kernels off and some edges
    at level 1 -- including every edge of one landmark and of one free pose.  Pinned: the index mapping of
    initializeOptimization (free poses first, landmarks second, ascending id, fixed / edge-less vertices excluded),
    computeActiveErrors touching active edges only (level-1 edges keep their phase-A _error), activeChi2 and
    activeRobustChi2, SparseOptimizer::update consuming the increment in index order, and the normal equations
    (Hpp, Hll, the 6x3 pose-landmark blocks, b) that constructQuadraticForm accumulates."""
    g = {k[len("graph_"):]: gold[k] for k in gold.files if k.startswith("graph_")}
    n_pose, n_point = len(g["pose"]), len(g["X"])
    prob = synth.Problem(g["pose"].copy(), g["fixed"].astype(np.uint8), np.tile(g["cam"], (n_pose, 1)), g["X"].copy(),
                         g["obs"][:, 0].astype(np.int32), g["obs"][:, 1].astype(np.int32), g["meas"].astype(np.float32))
    r = refba.RefBA(prob)
    # row A7: the normal equations of phase A at the initial estimates, assembled by the binary's own linearizeOplus +
    # constructQuadraticForm (Huber weighting through robustInformation) into blocks mapped the way BlockSolver maps them
    r.debug_phase(g["levA"], True, None)
    sysA = r.debug_system()
    for k in ("Hpp", "Hll", "Hpl", "b_pose", "b_point"):
        want = g[f"sysA_{k}"]
        np.testing.assert_allclose(sysA[k], want, rtol=0, atol=1e-13 * np.abs(want).max(), err_msg=k)
    assert not g["sysA_Hpp"][g["fixed"] == 1].any() and not g["sysA_b_pose"][g["fixed"] == 1].any()   # fixed: no blocks
    assert not g["sysA_Hpl"][g["fixed"][g["obs"][:, 0]] == 1].any()
    assert np.abs(g["sysA_Hpl"][g["fixed"][g["obs"][:, 0]] == 0]).min(axis=(1, 2)).max() > 0
    A = r.debug_phase(g["levA"], True, g["updA"])
    B = r.debug_phase(g["levB"], False, g["updB"])
    for tag, got in (("A", A), ("B", B)):
        assert np.array_equal(got["pose_index"], g[f"{tag}_pose_index"]), tag
        assert np.array_equal(got["point_index"], g[f"{tag}_point_index"]), tag
        np.testing.assert_allclose(got["err"], g[f"{tag}_err"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(got["chi2"], float(g[f"{tag}_chi2"]), rtol=1e-12)
        np.testing.assert_allclose(got["robust_chi2"], float(g[f"{tag}_robust_chi2"]), rtol=1e-12)
        np.testing.assert_allclose(got["poses"], g[f"{tag}_poses"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(got["points"], g[f"{tag}_points"], rtol=0, atol=1e-14)
    # what the fixture exercises
    assert g["A_robust_chi2"] < g["A_chi2"]                      # Huber's outlier branch was hit
    assert (g["A_pose_index"][g["fixed"] == 1] == -1).all() and (g["A_pose_index"][g["fixed"] == 0] >= 0).all()
    free_inactive = (g["B_pose_index"] == -1) & (g["fixed"] == 0)
    assert free_inactive.sum() == 1 and (g["B_point_index"] == -1).sum() >= 1
    lv1 = g["levB"] == 1
    assert lv1.sum() >= 5 and np.array_equal(g["B_err"][lv1], g["A_err"][lv1])      # stale _error of level-1 edges
    assert not np.array_equal(g["B_err"][~lv1], g["A_err"][~lv1])                  # active ones were recomputed
    # an excluded vertex is not moved by update()
    i = int(np.flatnonzero(free_inactive)[0])
    assert np.array_equal(g["B_poses"][i], g["A_poses"][i])
    assert not np.array_equal(g["A_poses"][1], g["pose"][1])


@pytest.mark.parametrize("case", [0, 1])
def test_levenberg_policy_matches_reference_binary(gold, synth, case):
    """Row A10/A11 against the binary: its own SparseOptimizer::optimize(30) + OptimizationAlgorithmLevenberg::solve ran
    on a g2o::Solver implemented in oracle/pin_libg2o_graph.py (normal equations from the binary's
    constructQuadraticForm, lambda on every diagonal, one dense solve), on two problems far from their optimum -- one
    with Huber kernels, one without.  The lambda handed to each trial is a complete fingerprint of lambda_0 = tau * max
    diag(H), every gain ratio rho = (chi - chi_trial) / (x.(lambda x + b) + 1e-3), the 1 - (2 rho - 1)^3 factor with
    its clamps, the lambda * nu / nu * 2 rule after a rejection and the accept / reject / stop decisions; the oracle's
    Schur + LDLT path must reproduce it trial by trial, stop after the same iteration and land on the same estimates."""
    g = {k[len(f"lm{case}_"):]: gold[k] for k in gold.files if k.startswith(f"lm{case}_")}
    n_pose = len(g["pose0"])
    prob = synth.Problem(g["pose0"].copy(), g["fixed"].astype(np.uint8), np.tile(g["cam"], (n_pose, 1)), g["X0"].copy(),
                         g["obs"][:, 0].astype(np.int32), g["obs"][:, 1].astype(np.int32), g["meas"].astype(np.float32))
    r = refba.RefBA(prob)
    r.solve_global(int(g["iters"]), bool(g["robust"]))
    tr = r.trace()
    lam = g["lambda"]
    assert len(tr) == len(lam), f"oracle made {len(tr)} trials, the binary {len(lam)}"
    # exact up to the point where a factor is unclamped; from there lambda carries the ~1e-7 reproducibility of the
    # stereo cost (float32 inverse depth) through rho
    np.testing.assert_allclose(tr[:, 3], lam, rtol=1e-6)
    first_unclamped = int(np.argmax(~(np.isclose(lam[1:] / lam[:-1], 1 / 3) | (lam[1:] > lam[:-1]) |
                                      np.isclose(lam[1:] / lam[:-1], 2 / 3)))) or len(lam)
    np.testing.assert_allclose(tr[:first_unclamped, 3], lam[:first_unclamped], rtol=1e-12)
    assert int(tr[:, 1].max()) + 1 == int(g["n_iterations"]) < int(g["iters"])      # both stop by the _nBad rule
    np.testing.assert_allclose(tr[0, 4], g["chi2"][0], rtol=1e-12)
    # the stereo residual rounds 1/z to float32 (types_six_dof_expmap.cpp:151): two exact solvers that differ by 1e-13
    # in the state can see that rounding flip, so cost and estimates are reproducible to ~1e-7 relative, not to 1e-12
    np.testing.assert_allclose(tr[tr[:, 7] == 1][-1, 5], g["chi2"][1], rtol=1e-7)
    # (and the toy problems stop by the _nBad rule on a weakly constrained valley: 1e-13 per step grows to ~1e-7 m)
    np.testing.assert_allclose(r.poses(), g["poses"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r.points(), g["points"], rtol=0, atol=1e-5)
    # what the fixture exercises: rejected trials (lambda grows), consecutive rejections (nu doubles), unclamped factors
    ratio = lam[1:] / lam[:-1]
    assert (ratio > 1).sum() >= 4 and (ratio > 7).sum() >= 1                      # nu = 2, 4, 8, ... in a row
    assert np.isclose(ratio, 1 / 3).any() and np.isclose(ratio, 2 / 3).any()      # both clamps of the factor
    if case == 1:
        assert ((ratio > 1 / 3 + 1e-6) & (ratio < 2 / 3 - 1e-6)).sum() >= 1      # and the cubic in between
    assert np.array_equal(tr[:-1, 7] == 0, ratio > 1)                             # a rejected trial <=> lambda grows


def test_two_pass_local_ba_matches_reference_binary(gold, synth):
    """The two-pass schedule of g2oOptimizer::LocalBundleAdjustment (g2oOptimizer.cc:923-976, 1119-1142) driven over the
    binary's own optimiser, edges, kernels and Levenberg (oracle/pin_libg2o_graph.py: make_lba): optimize(5) with Huber,
    chi2 / depth classification to level 1, kernels off, optimize(10), final flags.  The oracle's refba_solve_local must
    give the same lambda sequence in both passes, the same level-1 set, the same outlier flags and estimates."""
    g = {k[len("lba_"):]: gold[k] for k in gold.files if k.startswith("lba_")}
    n_pose = len(g["pose0"])
    prob = synth.Problem(g["pose0"].copy(), g["fixed"].astype(np.uint8), np.tile(g["cam"], (n_pose, 1)), g["X0"].copy(),
                         g["obs"][:, 0].astype(np.int32), g["obs"][:, 1].astype(np.int32), g["meas"].astype(np.float32))
    r = refba.RefBA(prob)
    r.solve_local(0)
    tr = r.trace()
    for p, key in ((0, "lambda1"), (1, "lambda2")):
        rows = tr[tr[:, 0] == p]
        assert len(rows) == len(g[key]), (p, len(rows), len(g[key]))
        np.testing.assert_allclose(rows[:, 3], g[key], rtol=1e-6)
        assert int(rows[:, 1].max()) + 1 == int(g["n_iterations"][p])
    assert np.array_equal(r.outliers(), g["outlier"])
    np.testing.assert_allclose(r.poses(), g["poses"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(r.points(), g["points"], rtol=0, atol=1e-7)
    # the fixture has rejected trials in pass 2, a real level-1 set, and every level-1 edge ends up flagged (its _error is
    # frozen at the pass-1 value)
    assert (g["lambda2"][1:] > g["lambda2"][:-1]).any()
    assert 10 < g["level2"].sum() < len(g["level2"]) // 2
    assert (g["outlier"][g["level2"] == 1] == 1).all()


def test_pose_only_optimisation_matches_reference_binary(gold):
    """g2oOptimizer::PoseOptimization (g2oOptimizer.cc:385-559, 655-690) driven over the binary's own VertexSE3Expmap,
    EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose objects, Huber kernels and Levenberg loop
    (oracle/pin_libg2o_graph.py: make_poseopt): four rounds of optimize(10) from the initial pose with the chi2
    re-classification in between.  The oracle's refba_pose_opt -- the checker of the pose-only CUDA kernel -- must make
    the same trials in every round with the same lambda, flag the same observations and end on the same pose."""
    g = {k[len("po_"):]: gold[k] for k in gold.files if k.startswith("po_")}
    pose, outlier, inliers, tr = refba.pose_opt(g["pose0"], g["cam"], g["xyz"], g["meas"])
    assert [int((tr[:, 0] == r).sum()) for r in range(4)] == list(g["round_trials"])
    assert [int(tr[tr[:, 0] == r][:, 1].max()) + 1 for r in range(4)] == list(g["n_iterations"])
    np.testing.assert_allclose(tr[:, 3], g["lambda"], rtol=1e-6)
    assert np.array_equal(outlier, g["outlier"]) and inliers == int(g["inliers"])
    np.testing.assert_allclose(pose, g["pose"], rtol=0, atol=1e-12)
    assert 20 < g["outlier"].sum() < 60 and (g["meas"][:, 2] < 0).any() and (g["meas"][:, 2] >= 0).any()
    assert (g["lambda"][1:] > g["lambda"][:-1]).sum() >= 4      # rejected trials inside the rounds


def test_sim3_arithmetic_matches_reference_binary(gold):
    """Row N3 groundwork: g2o::Sim3(Vector7d) (the exponential map with its four small-angle / small-scale branches,
    sim3.h:68-144), VertexSim3Expmap::oplusImpl with and without _fix_scale, and EdgeSim3::computeError = log(C * v1 *
    v2^-1) (operator*, inverse, log incl. the 3x3 LU solve) on real objects of the binary."""
    for u, want in zip(gold["sim3_upd"], gold["sim3_exp"]):
        np.testing.assert_allclose(refba.sim3_exp(u), want, rtol=0, atol=4e-15 * max(1.0, np.abs(want).max()))
    for s8, u, free, fix in zip(gold["sim3_exp"], gold["sim3_upd2"], gold["sim3_oplus_free"], gold["sim3_oplus_fix"]):
        np.testing.assert_allclose(refba.sim3_oplus(s8, u, False), free, rtol=0, atol=4e-15 * max(1.0, np.abs(free).max()))
        np.testing.assert_allclose(refba.sim3_oplus(s8, u, True), fix, rtol=0, atol=4e-15 * max(1.0, np.abs(fix).max()))
        assert fix[7] == s8[7] and free[7] != s8[7]                      # _fix_scale really pins the scale
    for m, a, b, want in zip(gold["sim3_meas"], gold["sim3_v1"], gold["sim3_v2"], gold["sim3_err"]):
        np.testing.assert_allclose(refba.sim3_edge_error(m, a, b), want, rtol=0, atol=1e-13 * max(1.0, np.abs(want).max()))
    th = np.linalg.norm(gold["sim3_upd"][:, :3], axis=1)
    sg = np.abs(gold["sim3_upd"][:, 6])
    for small_t in (True, False):                                        # all four branches of the exponential
        for small_s in (True, False):
            assert (((th < 1e-5) == small_t) & ((sg < 1e-5) == small_s)).sum() >= 5


@pytest.mark.parametrize("case", [0, 1])
def test_essential_graph_optimisation_matches_reference_binary(gold, case):
    """The optimisation g2oOptimizer::OptimizeEssentialGraph sets up (g2oOptimizer.cc:1212-1232, 1472-1478) run by the
    binary itself -- VertexSim3Expmap / EdgeSim3 objects, numeric Jacobians (central differences, delta 1e-9), identity
    information, Levenberg with setUserLambdaInit(1e-16), optimize(20) -- on a drifting loop with a closing edge, with
    fixed (stereo) and free (monocular) scale.  refba_pose_graph must make the same trials with the same lambda, stop in
    the same iteration (both end through ten rejected trials in a row, the qmax == 10 rule) and reach the same cost."""
    g = {k[len(f"pg{case}_"):]: gold[k] for k in gold.files if k.startswith(f"pg{case}_")}
    V, tr, done = refba.pose_graph(g["vert0"], g["fixed"], int(g["fix_scale"]), g["edges"], g["meas"], 20, 1e-16)
    assert done == int(g["n_iterations"]) and len(tr) == len(g["lambda"])
    np.testing.assert_allclose(tr[:, 3], g["lambda"], rtol=1e-6)
    assert tr[0, 3] == 1e-16
    np.testing.assert_allclose(tr[0, 4], g["chi2"][0], rtol=1e-12)
    np.testing.assert_allclose(tr[tr[:, 7] == 1][-1, 5], g["chi2"][1], rtol=1e-6)
    # numeric Jacobians with delta = 1e-9 carry ~1e-7 of noise per entry: the estimates agree to that, amplified
    np.testing.assert_allclose(V, g["vert"], rtol=0, atol=1e-4)
    assert g["chi2"][1] < 0.05 * g["chi2"][0]
    last = tr[tr[:, 1] == tr[-1, 1]]
    assert len(last) >= 10 and (last[:9, 7] == 0).all()                  # the run ends on the ten-trials rule


@pytest.mark.parametrize("case", [0, 1])
def test_optimize_sim3_matches_reference_binary(gold, case):
    """g2oOptimizer::OptimizeSim3 (g2oOptimizer.cc:1560-1796) run by the binary itself (oracle/pin_libg2o_graph.py:
    make_sim3opt) -- one VertexSim3Expmap carrying both cameras, fixed VertexSBAPointXYZ, EdgeSim3ProjectXYZ /
    EdgeInverseSim3ProjectXYZ with Huber kernels and their inherited NUMERIC Jacobians, Levenberg on a dense 7x7 --
    with fixed and free scale: optimize(5), chi2 test on the stored errors, optimize(10), final count.
    The oracle must reproduce both edges' errors and Jacobians at the initial estimate BIT FOR BIT, make the same
    trials with the same lambda while the decisions are above the noise of the 1e-9 differences, keep exactly the same
    matches and end on the same S12."""
    pre = f"s3o{case}_"
    g = {k[len(pre):]: gold[k] for k in gold.files if k.startswith(pre)}
    fs = bool(g["fix_scale"])
    e12, e21, J12, J21 = refba.sim3_match_linearize(g["s0"], g["cam8"], g["p1c"], g["p2c"], g["meas6"], fs)
    assert np.array_equal(e12, g["e12_0"]) and np.array_equal(e21, g["e21_0"])
    assert np.array_equal(J12, g["J12_0"]) and np.array_equal(J21, g["J21_0"])
    assert fs == (not J12[:, :, 6].any()) and np.abs(J12[:, :, :6]).max() > 100
    S, keep, n_in, tr = refba.optimize_sim3(g["s0"], g["cam8"], g["p1c"], g["p2c"], g["meas6"], float(g["th2"]), fs)
    assert np.array_equal(keep, g["keep"]) and n_in == int(g["nIn"])
    assert int(g["nBad"]) >= 5 and n_in >= 30                                  # the fixture drops matches in the first test
    np.testing.assert_allclose(S, g["s12"], rtol=0, atol=1e-6)
    # the first pass is far from convergence: identical trials; near the optimum the sign of a 1e-9-relative change of
    # the cost decides between accept / reject, in the binary as here
    n0 = int(g["n_trials_pass0"])
    assert int((tr[:, 0] == 0).sum()) >= min(n0, 4)
    m = min(n0, int((tr[:, 0] == 0).sum()), 4)
    np.testing.assert_allclose(tr[:m, 3], g["lambda"][:m], rtol=1e-6)
    k1 = int(np.argmax(tr[:, 0] == 1))
    np.testing.assert_allclose(tr[k1:k1 + 3, 3], g["lambda"][n0:n0 + 3], rtol=1e-6)   # lambda_0 of pass 2 and the next two
