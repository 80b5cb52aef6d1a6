"""The C++ adapter (sqrtlm-slam_b200/host/sqrtbaOptimizer.cc) that keeps the reference's
Optimizer::LocalBundleAdjustment / GlobalBundleAdjustemnt signatures (include/backend/Optimizer.h:50-56), driven the
way LocalMapping / LoopClosing drive it, against the oracle on the same map.  Map state is float32 (cv::Mat CV_32F in
the reference), so poses/points are compared after the same rounding."""
import ctypes

import numpy as np
import pytest

from oracle import refba

pytestmark = pytest.mark.gpu


def f32_pose_matrix(pose7, synth):
    R = synth.quat_to_rotmat(pose7[3:])
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R.astype(np.float32)
    T[:3, 3] = pose7[:3].astype(np.float32)
    return T


def test_local_ba_through_reference_api(pkg, synth):
    prob = synth.make_problem(31, 14, 5, 900, 6.0, stereo=True, name="adapter-lba")
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    free = np.nonzero(prob.pose_fixed == 0)[0]
    m.set_covisible(cur, [i for i in free if i != cur])      # local window = current KF + covisibles
    m.local_ba(cur)
    assert m.last_error() == ""
    ref = refba.RefBA(prob)
    ref.solve_local(0)
    P, X, F = ref.poses(), ref.points(), ref.outliers()
    for i in range(prob.n_pose):
        want = f32_pose_matrix(P[i], synth) if prob.pose_fixed[i] == 0 else f32_pose_matrix(prob.pose_qt[i], synth)
        np.testing.assert_allclose(m.pose(i), want, rtol=0, atol=2e-5)
    for j in range(prob.n_point):
        np.testing.assert_allclose(m.point(j), X[j].astype(np.float32), rtol=2e-6, atol=2e-5)
        assert m.point_updates(j) == 1                         # UpdateNormalAndDepth once per local point
    # outlier observations are erased both ways (g2oOptimizer.cc:1149-1161), inliers kept
    for k in range(prob.n_obs):
        kf, mp = int(prob.obs_pose[k]), int(prob.obs_point[k])
        assert m.has_observation(kf, mp) == (F[k] == 0)
        assert m.keyframe_sees(kf, mp) == (F[k] == 0)
    assert 0 < F.sum() < prob.n_obs


def test_local_ba_fork_defaults_through_reference_api(pkg, synth):
    """The adapter's DEFAULT behaviour is the reference fork's: monocular edges only (the stereo branch of its local BA
    is empty, g2oOptimizer.cc:914-916) and the 5 + 10 + 20 schedule (the third optimize(20) runs unconditionally,
    :1113-1114).  Oracle: the same window without the stereo observations, solve_local(20)."""
    from test_adapter_gather import mixed_mono_stereo
    prob = synth.make_problem(41, 14, 5, 1200, 6.0, stereo=True, name="adapter-fork")
    q, _ = mixed_mono_stereo(prob, seed=2, frac=0.6)
    m = pkg.host_harness.MockMap(q)
    cur = q.n_pose - 1
    free = np.nonzero(q.pose_fixed == 0)[0]
    m.set_covisible(cur, [i for i in free if i != cur])
    m.local_ba(cur, stereo_edges=False, two_pass=False)
    assert m.last_error() == ""
    keep = q.obs_meas[:, 2] < 0
    used = np.unique(q.obs_point[keep])
    remap = -np.ones(q.n_point, int)
    remap[used] = np.arange(len(used))
    sub = q.copy()
    sub.obs_pose, sub.obs_meas = q.obs_pose[keep], q.obs_meas[keep]
    sub.obs_point = remap[q.obs_point[keep]].astype(np.int32)
    sub.point_xyz = q.point_xyz[used]
    ref = refba.RefBA(sub)
    ref.solve_local(20)
    assert ref.trace()[:, 0].max() == 2
    P, X, F = ref.poses(), ref.points(), ref.outliers()
    for i in free:
        np.testing.assert_allclose(m.pose(int(i)), f32_pose_matrix(P[i], synth), rtol=0, atol=2e-5)
    for jj, j in enumerate(used):
        np.testing.assert_allclose(m.point(int(j)), X[jj].astype(np.float32), rtol=2e-6, atol=2e-5)
    for j in range(q.n_point):
        assert m.point_updates(j) == 1
        if remap[j] < 0:  # no monocular observation: a vertex without edges, never moved
            np.testing.assert_array_equal(m.point(j), q.point_xyz[j].astype(np.float32))
    ko = np.nonzero(keep)[0]
    for kk, k in enumerate(ko):
        kf, mp = int(q.obs_pose[k]), int(q.obs_point[k])
        assert m.has_observation(kf, mp) == (F[kk] == 0)
    for k in np.nonzero(~keep)[0][::5]:   # stereo observations are never erased: they had no edge
        assert m.has_observation(int(q.obs_pose[k]), int(q.obs_point[k]))


def test_local_ba_stop_flag_raised_leaves_map_untouched(pkg, synth):
    prob = synth.small_window(4)
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    m.set_covisible(cur, [i for i in np.nonzero(prob.pose_fixed == 0)[0] if i != cur])
    before = [m.pose(i).copy() for i in range(prob.n_pose)]
    flag = ctypes.c_bool(True)
    m.local_ba(cur, ctypes.byref(flag))
    for i in range(prob.n_pose):
        np.testing.assert_array_equal(m.pose(i), before[i])
    assert all(m.point_updates(j) == 0 for j in range(prob.n_point))


def test_local_ba_skips_bad_keyframes_and_points(pkg, synth):
    prob = synth.small_window(6, n_free=6, n_fixed=3, n_points=200)
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    m.set_covisible(cur, [i for i in np.nonzero(prob.pose_fixed == 0)[0] if i != cur])
    bad_kf, bad_mp = int(np.nonzero(prob.pose_fixed == 0)[0][0]), 5
    m.set_bad(kf=bad_kf, mp=bad_mp)
    p_before, x_before = m.pose(bad_kf).copy(), m.point(bad_mp).copy()
    m.local_ba(cur)
    np.testing.assert_array_equal(m.pose(bad_kf), p_before)   # bad KF is neither optimised nor written
    np.testing.assert_array_equal(m.point(bad_mp), x_before)
    assert m.point_updates(bad_mp) == 0
    # reference: the same problem without the bad keyframe's observations and the bad point
    keep = (prob.obs_pose != bad_kf) & (prob.obs_point != bad_mp)
    q = prob.copy()
    q.obs_pose, q.obs_point, q.obs_meas = prob.obs_pose[keep], prob.obs_point[keep], prob.obs_meas[keep]
    used = np.unique(q.obs_point)
    remap = -np.ones(prob.n_point, int)
    remap[used] = np.arange(len(used))
    q.obs_point = remap[q.obs_point].astype(np.int32)
    q.point_xyz = prob.point_xyz[used]
    q.pose_fixed = prob.pose_fixed.copy()
    q.pose_fixed[bad_kf] = 1
    ref = refba.RefBA(q)
    ref.solve_local(0)
    P = ref.poses()
    for i in np.nonzero(prob.pose_fixed == 0)[0]:
        if i != bad_kf:
            np.testing.assert_allclose(m.pose(int(i)), f32_pose_matrix(P[i], synth), rtol=0, atol=2e-5)


@pytest.mark.parametrize("n_loop_kf", [0, 7])
def test_global_ba_through_reference_api(pkg, synth, n_loop_kf):
    prob = synth.make_problem(17, 40, 1, 2000, 7.0, stereo=True, loop=True, cand_halfwidth=10, name="adapter-gba")
    m = pkg.host_harness.MockMap(prob)
    before = m.pose(3).copy()
    m.global_ba(10, False, n_loop_kf)
    ref = refba.RefBA(prob)
    ref.solve_global(10, False)
    P, X = ref.poses(), ref.points()
    gba = n_loop_kf != 0
    for i in range(prob.n_pose):
        np.testing.assert_allclose(m.pose(i, gba=gba), f32_pose_matrix(P[i], synth), rtol=0, atol=2e-5)
        if gba:   # staged result: mTcwGBA + marker, live pose untouched (g2oOptimizer.cc:324-329)
            assert m.gba_marker(i) == n_loop_kf
    if gba:
        np.testing.assert_array_equal(m.pose(3), before)
    for j in range(0, prob.n_point, 50):
        np.testing.assert_allclose(m.point(j, gba=gba), X[j].astype(np.float32), rtol=2e-6, atol=2e-5)


def test_pose_optimization_through_reference_api(pkg, synth):
    """sqrtbaOptimizer::PoseOptimization(Frame*) against the oracle: same inlier count, same mvbOutlier flags, the pose
    written back through Frame::SetPose (float32).  The fork only wires monocular edges (g2oOptimizer.cc:441-483): a
    keypoint with a right coordinate is ignored, an unmatched keypoint too."""
    pose0, cam, xyz, meas, _ = synth.frame_problem(seed=77, n_points=900, stereo=False, outlier_frac=0.1)
    meas = meas.copy()
    meas[5::40, 2] = 300.0                                    # a few keypoints with a right coordinate: skipped by the fork
    fr = pkg.host_harness.MockFrame(pose0, cam, xyz, meas)
    for i in (3, 17):
        fr.unmatch(i)
    used = np.ones(len(xyz), bool)
    used[5::40] = False
    used[[3, 17]] = False
    rp, rf, ri, _ = refba.pose_opt(pose0, cam, xyz[used], meas[used])
    inl = fr.pose_optimization()
    assert pkg.host_harness.lib().hh_last_error().decode() == ""
    assert inl == ri
    T, flags = fr.state()
    assert np.array_equal(flags[used], rf)
    assert not flags[~used].any()
    np.testing.assert_allclose(T, f32_pose_matrix(rp, synth), rtol=0, atol=2e-6)
    # batch entry point (relocalisation candidates): same result per frame
    frames = [pkg.host_harness.MockFrame(*synth.frame_problem(seed=80 + k, n_points=400 + 100 * k)[:4]) for k in range(3)]
    inl_b = pkg.host_harness.pose_optimization_batch(frames)
    for k, f in enumerate(frames):
        p0, c, x, m, _ = synth.frame_problem(seed=80 + k, n_points=400 + 100 * k)
        rp, rf, ri, _ = refba.pose_opt(p0, c, x, m)
        assert inl_b[k] == ri
        Tk, fk = f.state()
        assert np.array_equal(fk, rf)
        np.testing.assert_allclose(Tk, f32_pose_matrix(rp, synth), rtol=0, atol=2e-6)


@pytest.mark.parametrize("fix_scale", [True, False])
def test_essential_graph_through_reference_api(pkg, synth, fix_scale):
    """Optimizer::OptimizeEssentialGraph(pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections,
    bFixScale) (include/backend/Optimizer.h:62-67) on a mock map after a loop closure: keyframe poses become
    [R | t/s] of the optimised Sim3 (g2oOptimizer.cc:1459-1476), every good map point moves with its reference keyframe
    (:1479-1518).  Expected values: the oracle's pose-graph optimisation of the graph the adapter gathers."""
    import essential_graph_case as egc
    case = egc.build(pkg, synth, n_kf=48, seed=5, fix_scale=fix_scale)
    m = case["map"]
    args = (case["loop_kf"], case["cur_kf"], case["corrected"], case["non_corrected"], case["connections"], fix_scale)
    g = m.essential_graph(*args, optimise=False)
    Vr, trr, _ = refba.pose_graph(g["vert8"], g["fixed"], fix_scale, g["edge_ij"], g["meas8"], iters=20)
    assert trr[trr[:, 7] == 1][-1, 5] < 0.2 * trr[0, 4]
    m.essential_graph(*args, optimise=True)
    assert m.last_error() == ""
    scale = 1.0 + np.abs(Vr[:, 4:7]).max()
    moved = 0.0
    for k in range(case["n_kf"]):
        R = synth.quat_to_rotmat(Vr[k, :4] / np.linalg.norm(Vr[k, :4]))
        T = m.pose(k)
        np.testing.assert_allclose(T[:3, :3], R.astype(np.float32), rtol=0, atol=2e-5)
        np.testing.assert_allclose(T[:3, 3], (Vr[k, 4:7] / Vr[k, 7]).astype(np.float32), rtol=0, atol=2e-5 * scale)
        moved = max(moved, float(np.abs(T[:3, 3] - case["T"][k, :3, 3]).max()))
    assert moved > 0.005                                                 # the drift was actually distributed
    np.testing.assert_array_equal(m.pose(case["loop_kf"]), case["T"][case["loop_kf"]] if fix_scale else m.pose(case["loop_kf"]))
    for j in range(0, len(case["pts"]), 3):                               # Swr_corrected.map(Srw.map(P))
        r = int(case["ref"][j])
        X = case["pts"][j].astype(np.float64)
        c = synth.sim3_mul(g["vert8"][r], np.concatenate([[0, 0, 0, 1.0], X, [1.0]]))[4:7]
        w = synth.sim3_mul(synth.sim3_inv(Vr[r]), np.concatenate([[0, 0, 0, 1.0], c, [1.0]]))[4:7]
        np.testing.assert_allclose(m.point(j), w.astype(np.float32), rtol=0, atol=5e-5 * scale)
        assert m.point_updates(j) == 1


def test_global_ba_result_propagation_by_the_caller(pkg, synth):
    """Loop closing calls GlobalBundleAdjustemnt(map, 10, &stop, nLoopKF, false) and then spreads the STAGED result
    (mTcwGBA / mPosGBA / mnBAGlobalForKF, written by the adapter) over the spanning tree: keyframes and points created
    while the optimisation ran follow their parent / reference keyframe (LoopClosing.cc:987-1103; restated as caller
    code in host/harness.cc).  Checks that the adapter stages exactly what that code consumes."""
    prob = synth.make_problem(19, 30, 1, 1500, 7.0, stereo=True, loop=True, cand_halfwidth=8, name="gba-propagate")
    m = pkg.host_harness.MockMap(prob)
    for k in range(1, prob.n_pose):
        m.set_parent(k, k - 1)
    m.set_origin(0)
    ref_kf = np.full(prob.n_point, -1)
    for o in range(prob.n_obs):
        if ref_kf[prob.obs_point[o]] < 0:
            ref_kf[prob.obs_point[o]] = prob.obs_pose[o]
    for j in range(prob.n_point):
        m.set_ref_kf(j, int(ref_kf[j]))
    n_loop = 9
    m.global_ba(10, False, n_loop)
    late_kf, late_mp = prob.n_pose - 1, int(np.nonzero(ref_kf == 5)[0][0])
    m.forget_gba(kf=late_kf, mp=late_mp)
    before = [m.pose(i).copy() for i in range(prob.n_pose)]
    staged = [m.pose(i, gba=True).copy() for i in range(prob.n_pose)]
    x_before = m.point(late_mp).copy()
    m.apply_gba(n_loop)
    ref = refba.RefBA(prob)
    ref.solve_global(10, False)
    P, X = ref.poses(), ref.points()
    for i in range(prob.n_pose - 1):
        np.testing.assert_array_equal(m.pose(i), staged[i])
        np.testing.assert_allclose(m.pose(i), f32_pose_matrix(P[i], synth), rtol=0, atol=2e-5)
    # the late keyframe keeps its pose RELATIVE to its parent: Tchild * Twc_parent(before) * TcwGBA_parent
    par = late_kf - 1
    want = before[late_kf].astype(np.float64) @ np.linalg.inv(before[par].astype(np.float64)) @ staged[par].astype(np.float64)
    np.testing.assert_allclose(m.pose(late_kf), want, rtol=0, atol=5e-5)
    assert m.gba_marker(late_kf) == n_loop
    for j in range(0, prob.n_point, 40):
        if j != late_mp:
            np.testing.assert_allclose(m.point(j), X[j].astype(np.float32), rtol=2e-6, atol=2e-5)
    # the late point moves with its reference keyframe: Twc_after * (Tcw_before * X)
    r = int(ref_kf[late_mp])
    Xc = before[r][:3, :3].astype(np.float64) @ x_before.astype(np.float64) + before[r][:3, 3]
    Ta = np.linalg.inv(m.pose(r).astype(np.float64))
    np.testing.assert_allclose(m.point(late_mp), Ta[:3, :3] @ Xc + Ta[:3, 3], rtol=0, atol=5e-5)


@pytest.mark.parametrize("fix_scale", [False, True])
def test_optimize_sim3_through_reference_api(pkg, synth, fix_scale):
    """Optimizer::OptimizeSim3(pKF1, pKF2, vpMatches1, g2oS12, th2, bFixScale) (include/backend/Optimizer.h:68-69) as
    LoopClosing::ComputeSim3 calls it (LoopClosing.cc:513): returns nIn, writes the refined S12 and sets the matches it
    rejects to NULL (g2oOptimizer.cc:1733, 1773).  Expected values: the oracle on the arrays the adapter gathers."""
    case = synth.sim3_pair(seed=8, n_matches=120, fix_scale=fix_scale)
    m, match = pkg.host_harness.MockMap.sim3_candidates(case, seed=2)
    match = match.copy()
    match[[3, 50]] = -1
    g = m.gather_sim3(0, 1, match)
    So, keep_o, nin_o, _ = refba.optimize_sim3(case[0], g["cam8"], g["p1c"], g["p2c"], g["meas6"], 10.0, fix_scale)
    n_in, after, S = m.optimize_sim3(0, 1, match, case[0], 10.0, fix_scale)
    assert m.last_error() == ""
    assert n_in == nin_o and 60 < n_in < 118
    want = match.copy()
    want[g["index"][keep_o == 0]] = -1
    assert np.array_equal(after, want) and (after == -1).sum() > 5
    sgn = np.sign(np.sum(S[:4] * So[:4]))
    np.testing.assert_allclose(S[:4] * sgn, So[:4], rtol=0, atol=1e-6)
    np.testing.assert_allclose(S[4:], So[4:], rtol=0, atol=1e-5)
    # too few matches: the reference returns 0 and leaves g2oS12 alone, but has already dropped what failed the first test
    few = np.full(len(match), -1, np.int32)
    few[:8] = match[:8] if (match[:8] >= 0).all() else match[5:13]
    n_in2, after2, S2 = m.optimize_sim3(0, 1, few, case[0], 10.0, fix_scale)
    assert n_in2 == 0 and np.array_equal(S2, case[0])
