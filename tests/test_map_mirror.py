"""Incremental observation mirror of the map (host/map_mirror.h): the adapter's gather
reads flat per-point lists that the map's mutation sites keep current instead of copying MapPoint::GetObservations()
(a std::map copied under the point's mutex, MapPoint.cc:212-215) for every point of every call.  CPU only."""
import hashlib

import numpy as np
import pytest


def digest(flat):
    return hashlib.sha1(b"".join(np.ascontiguousarray(flat[k]).tobytes() for k in sorted(flat))).hexdigest()


@pytest.fixture()
def window(pkg, synth):
    prob = synth.make_problem(41, 16, 5, 2500, 7.0, stereo=True, name="mirror")
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    m.set_covisible(cur, [int(i) for i in np.nonzero(prob.pose_fixed == 0)[0] if i != cur])
    yield prob, m, cur
    m.mirror_detach()


def test_gather_through_the_mirror_is_bit_identical(window, pkg):
    prob, m, cur = window
    plain = m.gather(cur)
    assert pkg.host_harness.MockMap.mirror_points() == 0     # off: nothing is tracked, the map copies are used
    m.mirror_attach()
    assert pkg.host_harness.MockMap.mirror_points() == prob.n_point and m.mirror_mismatches() == 0
    mirrored = m.gather(cur)
    assert digest(plain) == digest(mirrored)
    assert digest(m.gather(-1)) == digest(m.gather(-1))
    m.mirror_detach()
    assert digest(m.gather(cur)) == digest(plain)


def test_mirror_follows_the_map_through_its_mutation_sites(window):
    prob, m, cur = window
    m.mirror_attach()
    rng = np.random.default_rng(0)
    # erase a tenth of the observations, add new ones, retire some points -- all through MapPoint's own methods
    for k in rng.choice(prob.n_obs, prob.n_obs // 10, replace=False):
        m.erase_observation(int(prob.obs_pose[k]), int(prob.obs_point[k]))
    for _ in range(300):
        m.add_observation(int(rng.integers(prob.n_pose)), int(rng.integers(prob.n_point)), rng.uniform(0, 1200), rng.uniform(0, 370),
                          ur=-1.0 if rng.random() < 0.5 else rng.uniform(0, 1200), octave=int(rng.integers(8)))
    for mp in rng.choice(prob.n_point, 40, replace=False):
        m.set_point_bad(int(mp))
    assert m.mirror_mismatches() == 0
    with_mirror = m.gather(cur)
    m.mirror_detach()
    assert digest(m.gather(cur)) == digest(with_mirror)      # the map copies tell the same story
    assert len(with_mirror["obs_pose"]) < prob.n_obs          # ... and it is a different window than before


def test_concurrent_writers_and_snapshots(window):
    prob, m, cur = window
    m.mirror_attach()
    snaps = m.mirror_stress(n_writers=4, rounds=3)
    assert snaps >= 1
    assert m.mirror_mismatches() == 0
    a = m.gather(cur)
    m.mirror_detach()
    assert digest(m.gather(cur)) == digest(a)


def test_unknown_point_falls_back_to_the_map(pkg, synth):
    # a map the mirror was never attached to while the mirror is on for another one: Snapshot() refuses, the gather is right
    prob = synth.make_problem(43, 10, 4, 600, 6.0, stereo=True)
    a = pkg.host_harness.MockMap(prob)
    b = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    for mm in (a, b):
        mm.set_covisible(cur, [int(i) for i in np.nonzero(prob.pose_fixed == 0)[0] if i != cur])
    a.mirror_attach()
    try:
        assert b.mirror_mismatches() == -1
        assert digest(b.gather(cur)) == digest(a.gather(cur))
    finally:
        a.mirror_detach()


def test_gather_cost_with_and_without_the_mirror(pkg, synth, capsys):
    prob = synth.config_c0(0)
    m = pkg.host_harness.MockMap(prob)
    cur = prob.n_pose - 1
    m.set_covisible(cur, [int(i) for i in np.nonzero(prob.pose_fixed == 0)[0] if i != cur])
    t_map = m.time_gather(cur, 7)
    m.mirror_attach()
    try:
        t_mirror = m.time_gather(cur, 7)
    finally:
        m.mirror_detach()
    with capsys.disabled():
        print(f"\n[map mirror] local-window gather of a C0-shaped map: {t_map / 1e3:.2f} ms from map copies, "
              f"{t_mirror / 1e3:.2f} ms from the mirror")
    assert t_mirror < 2.5 * t_map   # a sanity bound only (shared CI hosts are noisy); the measured figures are printed above
