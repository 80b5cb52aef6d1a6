"""The host half of the adapter (sqrtlm-slam_b200/host/sqrtbaOptimizer.cc) on CPU: window selection and gather of
Optimizer::LocalBundleAdjustment (g2oOptimizer.cc:709-780, 805-919) and the gather of BundleAdjustment (:137-295), from a
map of header-compatible KeyFrame / MapPoint objects to the flat arrays of sqrtba_set_problem -- without solving, so no
GPU is needed.  The solve itself is covered by tests/test_gpu_adapter.py."""
import os

import numpy as np
import pytest


def expected_pose_qt(prob, synth):
    """What the adapter reads: float32 Tcw matrices converted as Converter::toSE3Quat does."""
    Rf = synth.quat_to_rotmat(prob.pose_qt[:, 3:]).astype(np.float32).astype(np.float64)
    t = prob.pose_qt[:, :3].astype(np.float32).astype(np.float64)
    return np.concatenate([t, synth.rotmat_to_quat_eigen(Rf)], axis=1)


def check_against_problem(flat, prob, synth, fixed):
    assert np.array_equal(flat["kf_ids"], np.arange(prob.n_pose))
    assert np.array_equal(flat["mp_ids"], np.arange(prob.n_point))
    np.testing.assert_allclose(flat["pose_qt"], expected_pose_qt(prob, synth), rtol=0, atol=1e-15)
    assert np.array_equal(flat["pose_fixed"], fixed)
    np.testing.assert_allclose(flat["cam"], np.tile(prob.cam[0].astype(np.float32).astype(np.float64), (prob.n_pose, 1)))
    assert np.array_equal(flat["point_xyz"], prob.point_xyz.astype(np.float32).astype(np.float64))
    # observations: grouped by landmark in ascending id, pose-sorted inside -- exactly the generator's order
    assert np.array_equal(flat["obs_point"], prob.obs_point)
    assert np.array_equal(flat["obs_pose"], prob.obs_pose)
    assert np.array_equal(flat["obs_meas"], prob.obs_meas)


def test_local_window_gather(pkg, synth):
    prob = synth.make_problem(31, 14, 5, 900, 6.0, stereo=True, name="gather-lba")
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    free = np.nonzero(prob.pose_fixed == 0)[0]
    m.set_covisible(cur, [int(i) for i in free if i != cur])
    flat = m.gather(cur)
    # local keyframes are free except mnId 0 (g2oOptimizer.cc:813), every other observer of a local point is fixed
    check_against_problem(flat, prob, synth, prob.pose_fixed)
    assert 0 < flat["pose_fixed"].sum() < prob.n_pose


def test_gather_is_thread_count_invariant(pkg, synth, monkeypatch):
    import subprocess
    import sys
    code = (
        "import importlib, sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "pkg = importlib.import_module('sqrtlm-slam_b200')\n"
        "prob = pkg.synth.make_problem(7, 12, 4, 3000, 7.0, stereo=True)\n"
        "m = pkg.host_harness.MockMap(prob)\n"
        "cur = prob.n_pose - 1\n"
        "m.set_covisible(cur, [int(i) for i in np.nonzero(prob.pose_fixed == 0)[0] if i != cur])\n"
        "f = m.gather(cur)\n"
        "import hashlib\n"
        "print(hashlib.sha1(b''.join(np.ascontiguousarray(f[k]).tobytes() for k in sorted(f))).hexdigest())\n")
    digests = set()
    for t in ("1", "3", "8"):
        env = dict(os.environ, SQRTBA_ADAPTER_THREADS=t)
        digests.add(subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.strip())
    assert len(digests) == 1, digests


def test_local_window_skips_bad_and_unrelated(pkg, synth):
    prob = synth.small_window(6, n_free=6, n_fixed=3, n_points=200)
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    free = [int(i) for i in np.nonzero(prob.pose_fixed == 0)[0]]
    bad_kf, bad_mp = free[0], 5
    m.set_covisible(cur, [i for i in free if i != cur])
    m.set_bad(kf=bad_kf, mp=bad_mp)
    flat = m.gather(cur)
    assert bad_kf not in flat["kf_ids"] and bad_mp not in flat["mp_ids"]
    local_kfs = [i for i in free if i != bad_kf]
    local_pts = np.unique(prob.obs_point[np.isin(prob.obs_pose, local_kfs)])   # seen by a (good) local keyframe
    keep = (prob.obs_pose != bad_kf) & (prob.obs_point != bad_mp) & np.isin(prob.obs_point, local_pts)
    kf_of = {int(k): i for i, k in enumerate(flat["kf_ids"])}
    mp_of = {int(k): i for i, k in enumerate(flat["mp_ids"])}
    # every surviving observation is there, in the same order, with compacted indices
    want_pose = np.array([kf_of[int(p)] for p in prob.obs_pose[keep]])
    want_point = np.array([mp_of[int(p)] for p in prob.obs_point[keep]])
    assert np.array_equal(flat["obs_pose"], want_pose) and np.array_equal(flat["obs_point"], want_point)
    assert np.array_equal(flat["obs_meas"], prob.obs_meas[keep])
    # a window made of only two keyframes: the other observers of their points become fixed, unrelated ones drop out
    m2 = pkg.host_harness.MockMap(prob)
    m2.set_covisible(cur, [free[-2]])
    f2 = m2.gather(cur)
    loc = {cur, free[-2]}
    fixed_ids = set(int(k) for k, f in zip(f2["kf_ids"], f2["pose_fixed"]) if f)
    assert loc.isdisjoint(fixed_ids - {0}) and set(int(k) for k in f2["kf_ids"]) >= loc
    seen_by_local = np.isin(prob.obs_pose, list(loc))
    assert set(int(p) for p in f2["mp_ids"]) == set(int(p) for p in np.unique(prob.obs_point[seen_by_local]))


def test_global_gather(pkg, synth):
    prob = synth.make_problem(11, 10, 1, 500, 5.0, stereo=True, name="gather-gba")
    m = pkg.host_harness.MockMap(prob)
    flat = m.gather(-1)
    fixed = np.zeros(prob.n_pose, np.uint8)
    fixed[0] = 1  # only mnId 0 (g2oOptimizer.cc:154)
    check_against_problem(flat, prob, synth, fixed)


def test_local_write_back(pkg, synth):
    """Outlier erasure both ways, SetPose for local keyframes only (float32 Tcw), SetWorldPos + one UpdateNormalAndDepth
    per local point (g2oOptimizer.cc:1145-1189) -- the parallel write-back must do exactly that."""
    prob = synth.make_problem(5, 12, 4, 2500, 6.0, stereo=True, name="write-back")   # enough points for several threads
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    free = np.nonzero(prob.pose_fixed == 0)[0]
    m.set_covisible(cur, [int(i) for i in free if i != cur])
    flat = m.gather(cur)
    rng = np.random.default_rng(0)
    P = flat["pose_qt"].copy()
    P[:, :3] += rng.normal(0, 0.1, (len(P), 3))
    X = flat["point_xyz"] + rng.normal(0, 0.05, flat["point_xyz"].shape)
    F = (rng.random(len(flat["obs_pose"])) < 0.1).astype(np.uint8)
    # a fresh map: the window-selection markers (mnBALocalForKF) of the gather above would hide the points from a second
    # selection with the same keyframe id, exactly as in the reference
    m = pkg.host_harness.MockMap(prob)
    m.set_covisible(cur, [int(i) for i in free if i != cur])
    before = [m.pose(i).copy() for i in range(prob.n_pose)]
    m.apply_local(cur, P, X, F)
    for i in range(prob.n_pose):
        if prob.pose_fixed[i]:
            np.testing.assert_array_equal(m.pose(i), before[i])          # fixed keyframes are not written
        else:
            np.testing.assert_allclose(m.pose(i)[:3, 3], P[i, :3].astype(np.float32), rtol=0, atol=0)
            np.testing.assert_allclose(m.pose(i)[:3, :3], synth.quat_to_rotmat(P[i, 3:]).astype(np.float32), rtol=0, atol=1e-7)
    for j in range(prob.n_point):
        np.testing.assert_array_equal(m.point(j), X[j].astype(np.float32))
        assert m.point_updates(j) == 1
    for k in range(prob.n_obs):
        kf, mp = int(prob.obs_pose[k]), int(prob.obs_point[k])
        assert m.has_observation(kf, mp) == (F[k] == 0) and m.keyframe_sees(kf, mp) == (F[k] == 0)


def mixed_mono_stereo(prob, seed=0, frac=0.5):
    """Every `frac` of the points loses the right coordinate in ALL its observations (monocular points), the others stay
    stereo -- so the fork's local BA (monocular edges only) sees whole points drop out."""
    q = prob.copy()
    rng = np.random.default_rng(seed)
    mono_pt = rng.random(prob.n_point) < frac
    q.obs_meas = prob.obs_meas.copy()
    q.obs_meas[mono_pt[prob.obs_point], 2] = -1.0
    return q, mono_pt


def test_fork_default_drops_stereo_observations(pkg, synth):
    """The reference fork's local BA leaves the stereo branch empty (g2oOptimizer.cc:914-916): the adapter's default
    (options().local_ba_stereo_edges = false) hands over the monocular observations only; points without one stay in
    the map, keep their position and still get their UpdateNormalAndDepth at write-back (:1180-1188)."""
    prob = synth.make_problem(11, 10, 3, 600, 5.0, stereo=True, name="fork-default")
    q, mono_pt = mixed_mono_stereo(prob, seed=1)
    m = pkg.host_harness.MockMap(q)
    cur = q.n_pose - 1
    free = np.nonzero(q.pose_fixed == 0)[0]
    m.set_covisible(cur, [int(i) for i in free if i != cur])
    flat = m.gather(cur, stereo_edges=False)
    keep = q.obs_meas[:, 2] < 0
    assert 0 < keep.sum() < q.n_obs
    assert len(flat["obs_pose"]) == keep.sum() and (flat["obs_meas"][:, 2] < 0).all()
    used = np.unique(q.obs_point[keep])
    assert np.array_equal(flat["mp_ids"], used)
    np.testing.assert_array_equal(flat["obs_meas"], q.obs_meas[keep])
    np.testing.assert_array_equal(flat["mp_ids"][flat["obs_point"]], q.obs_point[keep])
    # write-back: every local point is visited once, the edge-less ones keep their position
    m = pkg.host_harness.MockMap(q)
    m.set_covisible(cur, [int(i) for i in free if i != cur])
    X = flat["point_xyz"] + 0.01
    before = [m.point(j).copy() for j in range(q.n_point)]
    m.apply_local(cur, flat["pose_qt"], X, np.zeros(len(flat["obs_pose"]), np.uint8), stereo_edges=False)
    used_set = set(int(u) for u in used)
    for j in range(q.n_point):
        assert m.point_updates(j) == 1
        if j in used_set:
            np.testing.assert_array_equal(m.point(j), X[list(used).index(j)].astype(np.float32))
        else:
            np.testing.assert_array_equal(m.point(j), before[j])


def test_essential_graph_gather(pkg, synth):
    """The graph sqrtbaOptimizer::OptimizeEssentialGraph builds from the map (no GPU): vertices = Sim3-corrected poses
    where loop closing has them, the keyframe's own pose otherwise; edges by the reference's rules
    (g2oOptimizer.cc:1306-1448): new loop connections (weight >= 100 except the loop pair itself), spanning tree,
    earlier loop edges, covisibility >= 100 to older keyframes unless already inserted; relative poses from the
    NON-corrected estimates."""
    import essential_graph_case as egc
    case = egc.build(pkg, synth)
    g = case["map"].essential_graph(case["loop_kf"], case["cur_kf"], case["corrected"], case["non_corrected"],
                                    case["connections"], True, optimise=False)
    vert, edges, meas = egc.expected_graph(case, synth)
    assert g["present"].all() and g["fixed"].sum() == 1 and g["fixed"][case["loop_kf"]] == 1
    np.testing.assert_allclose(g["vert8"], vert, rtol=0, atol=1e-12)
    got = {(int(i), int(j)): mm for (i, j), mm in zip(g["edge_ij"], g["meas8"])}
    want = {e: mm for e, mm in zip(edges, meas)}
    assert len(got) == len(g["edge_ij"]) and set(got) == set(want)
    for e in want:
        np.testing.assert_allclose(got[e], want[e], rtol=0, atol=1e-12)
    # spot checks of the rules
    cur, n = case["cur_kf"], case["n_kf"]
    assert (cur, 0) in got and (cur, 1) in got and (cur - 1, 1) in got and (cur - 2, 2) not in got
    assert (12, 3) in got and (4, 2) in got and (5, 3) not in got and (6, 3) not in got


def test_sim3_gather(pkg, synth):
    """The arrays sqrtbaOptimizer::OptimizeSim3 builds from the two keyframes (no GPU), by the reference's rules
    (g2oOptimizer.cc:1622-1712): a match is used when both map points are good and the second is observed in pKF2; the
    points are moved into their own keyframe's frame on the FLOAT matrices (R * X + t as one cv::Mat GEMM: double
    accumulator, one rounding); keypoints and level sigmas come from keypoint i of pKF1 and from the keypoint of pKF2
    that observes the matched point; both cameras' mK feed the vertex."""
    from importlib import import_module
    hh = import_module(pkg.__name__ + ".host_harness")
    case = synth.sim3_pair(seed=3, n_matches=60)
    cam2 = (700.0, 705.0, 600.5, 180.25)
    m, match = hh.MockMap.sim3_candidates(case, seed=1, cam2=cam2)
    n = len(match)
    match = match.copy()
    match[[4, 17]] = -1                       # no match for these keypoints
    m.set_bad(mp=9)                           # pMP1 bad
    m.set_bad(mp=n + 30)                      # pMP2 bad
    match[40] = 12                            # a point that pKF2 does not observe: GetIndexInKeyFrame < 0
    g = m.gather_sim3(0, 1, match)
    used = [i for i in range(n) if i not in (4, 17, 9, 30, 40)]
    assert list(g["index"]) == used
    np.testing.assert_array_equal(g["cam8"], np.concatenate([np.float32(case[1][:4]), np.float32(cam2)]).astype(np.float64))
    T, Xw = m.Tcw.astype(np.float64), m.Xw.astype(np.float64)

    def to_cam(k, X):
        return (T[k, :3, :3] @ X + T[k, :3, 3]).astype(np.float32).astype(np.float64)

    for row, i in enumerate(used):
        np.testing.assert_array_equal(g["p1c"][row], to_cam(0, Xw[i]))
        np.testing.assert_array_equal(g["p2c"][row], to_cam(1, Xw[n + i]))
    np.testing.assert_array_equal(g["meas6"], case[4][used])
    # the float round trip through world coordinates keeps the points within float precision of the generator's
    np.testing.assert_allclose(g["p1c"], case[2][used], rtol=0, atol=2e-4)
