"""Host-side plan of sqrtba_set_problem (items -> tiles -> run tables -> landmark order), checked WITHOUT a GPU.

The plan is what the kernels index by: a wrong rank, run boundary or column count silently
corrupts the pose-side reduction, so every structural invariant the kernels rely on is asserted here on problems of
all three shapes: batched small windows, a single window, and a big window (> 128 free poses, internally re-ordered).
sqrtba_debug_plan runs the very same set_problem code in plan-only mode (no CUDA call is made).
"""
import importlib

import numpy as np
import pytest

capi = importlib.import_module("sqrtlm-slam_b200.capi")
synth = importlib.import_module("sqrtlm-slam_b200.synth")

T = {k: i for i, k in enumerate(capi.TILE_COLS)}
JQ_HDR, JQ_ROWS = 8, 14  # sqrtba_kernels.cuh: 4 geometry rows {x/z, y/z, 1/z, w} + 9 rows of Q1 + the 8-byte meta entries


def internal_order(prob, plan):
    """Observation arrays in the solver's internal order (landmarks permuted, their observations kept contiguous)."""
    order = plan["landmark_order"]
    if not plan["reordered"]:
        assert np.array_equal(order, np.arange(prob.n_point))
        return prob.obs_pose.copy(), prob.obs_point.copy()
    first = np.searchsorted(prob.obs_point, np.arange(prob.n_point), side="left")
    last = np.searchsorted(prob.obs_point, np.arange(prob.n_point), side="right")
    idx = np.concatenate([np.arange(first[l], last[l]) for l in order])
    new_id = np.empty(prob.n_point, np.int64)
    new_id[order] = np.arange(prob.n_point)
    return prob.obs_pose[idx], new_id[prob.obs_point[idx]]


def free_slots(prob, pose_ptr):
    """Window-relative slot of every pose (-1 for fixed poses): free poses numbered in order inside their window."""
    slot = np.full(prob.n_pose, -1, np.int64)
    for w in range(len(pose_ptr) - 1):
        a, b = int(pose_ptr[w]), int(pose_ptr[w + 1])
        fr = np.flatnonzero(prob.pose_fixed[a:b] == 0)
        slot[a + fr] = np.arange(len(fr))
    return slot


def check_plan(prob, plan, pose_ptr, obs_ptr):
    tiles = plan["tiles"]
    obs_pose, obs_point = internal_order(prob, plan)
    slot = free_slots(prob, pose_ptr)
    oslot = slot[obs_pose]
    n_tile = plan["n_tile"]
    assert n_tile == len(tiles) and n_tile > 0

    # tiles partition the observations, in order, and never straddle a window
    assert tiles[0, T["o0"]] == 0 and tiles[-1, T["o1"]] == prob.n_obs
    assert np.array_equal(tiles[1:, T["o0"]], tiles[:-1, T["o1"]])
    win_of_obs = np.searchsorted(np.asarray(obs_ptr), np.arange(prob.n_obs), side="right") - 1
    # items are consecutive and cover all items
    assert tiles[0, T["item0"]] == 0
    assert np.array_equal(tiles[1:, T["item0"]], tiles[:-1, T["item0"]] + tiles[:-1, T["nitem"]])
    assert tiles[-1, T["item0"]] + tiles[-1, T["nitem"]] == plan["n_item"]

    jq = 0
    run_ptr, runs = plan["tile_run_ptr"], plan["tile_runs"]
    assert run_ptr[0] == 0 and run_ptr[-1] == plan["n_run_ints"]
    for t in range(n_tile):
        ti = tiles[t]
        o0, o1 = int(ti[T["o0"]]), int(ti[T["o1"]])
        cnt = ti[T["cnt0"]:T["cnt0"] + 4]
        fcnt = ti[T["fcnt0"]:T["fcnt0"] + 4]
        assert np.all(win_of_obs[o0:o1] == ti[T["win"]])
        assert 1 <= ti[T["nitem"]] <= 4
        assert cnt.sum() == o1 - o0
        assert np.all(cnt[ti[T["nitem"]]:] == 0) and np.all(cnt[:ti[T["nitem"]]] > 0)
        # a landmark never straddles items: item boundaries are landmark boundaries
        bounds = o0 + np.concatenate([[0], np.cumsum(cnt[:ti[T["nitem"]]])])
        for b in bounds[1:-1]:
            assert obs_point[b - 1] != obs_point[b]
        if o0 > 0:
            assert obs_point[o0 - 1] != obs_point[o0]
        # JQ block geometry (one TMA bulk copy per tile: 16-byte aligned offset and size)
        jq_off = int(ti[T["jq_off_lo"]]) + (int(ti[T["jq_off_hi"]]) << 31)
        assert jq_off == jq + JQ_HDR
        assert ti[T["blk_doubles"]] % 2 == 0 and jq_off % 2 == 0 and ti[T["nt"]] % 2 == 0
        jq += int(ti[T["blk_doubles"]])
        n_run_ints = run_ptr[t + 1] - run_ptr[t]
        if ti[T["is_long"]]:
            assert ti[T["nitem"]] == 1 and cnt[0] > 32
            assert len(np.unique(obs_point[o0:o1])) == 1
            assert ti[T["nt"]] == (o1 - o0 + 1) // 2 * 2
            assert n_run_ints == 0
            assert ti[T["blk_doubles"]] == JQ_HDR + JQ_ROWS * ti[T["nt"]]
            continue
        assert np.all(cnt <= 32)
        fr = oslot[o0:o1] >= 0
        # per-item free counts and the column count
        for i in range(ti[T["nitem"]]):
            assert fcnt[i] == fr[bounds[i] - o0:bounds[i + 1] - o0].sum()
        assert ti[T["nfree"]] == fr.sum() == fcnt.sum()
        assert ti[T["nt"]] == (ti[T["nfree"]] + 1) // 2 * 2
        # run table: nrun+1 rank offsets then nrun slots, slots strictly increasing, runs cover ranks [0, nfree)
        nrun = int(ti[T["nrun"]])
        assert n_run_ints == 2 * nrun + 1
        assert ti[T["blk_doubles"]] == JQ_HDR + JQ_ROWS * ti[T["nt"]] + 2 * ((n_run_ints + 3) // 4)
        tab = runs[run_ptr[t]:run_ptr[t + 1]]
        roff, rslot = tab[:nrun + 1], tab[nrun + 1:]
        assert roff[0] == 0 and roff[-1] == ti[T["nfree"]]
        assert np.all(np.diff(roff) > 0)
        assert np.all(np.diff(rslot) > 0)
        assert np.array_equal(rslot, np.unique(oslot[o0:o1][fr]))
        # meta words: low half = window-relative slot, high half = rank; ranks are a bijection onto [0, nfree) and
        # every rank falls inside the run of its slot
        lp = plan["obs_lp"][o0:o1][fr]
        lo, rank = (lp & 0xFFFF).astype(np.int64), ((lp >> 16) & 0x7FFF).astype(np.int64)   # bit 31: stereo edge
        assert np.array_equal(lo, oslot[o0:o1][fr])
        assert np.array_equal(np.sort(rank), np.arange(ti[T["nfree"]]))
        run_of_rank = np.searchsorted(roff, rank, side="right") - 1
        assert np.array_equal(rslot[run_of_rank], lo)
        # inside one run, ranks follow observation order (the reduction order is therefore fixed -> reproducible)
        for r in range(nrun):
            rr = rank[run_of_rank == r]
            assert np.all(np.diff(rr) == 1) and rr[0] == roff[r]
    assert jq == plan["jq_doubles"]


def test_plan_single_small_window():
    w = synth.small_window(seed=3)
    plan = capi.debug_plan(w)
    assert plan["smallwin"] and plan["pq_shared"] and not plan["reordered"]
    check_plan(w, plan, [0, w.n_pose], [0, w.n_obs])


def test_plan_batch_of_windows_is_thread_count_invariant():
    wins = [synth.small_window(seed=s, n_free=3 + s % 5, n_points=60 + 17 * s) for s in range(9)]
    prob, pp, tp, op = synth.concat_windows(wins)
    plan = capi.debug_plan(prob, pp, tp, op, host_threads=1)
    check_plan(prob, plan, pp, op)
    # the multi-threaded builder must produce the identical plan (chunks are merged in order)
    plan4 = capi.debug_plan(prob, pp, tp, op, host_threads=4)
    for k, v in plan.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, plan4[k]), k
        else:
            assert v == plan4[k], k
    # every window's tiles only hold that window's observations
    assert set(plan["tiles"][:, T["win"]]) == set(range(len(wins)))


def test_plan_long_tracks():
    # a few landmarks seen by more than 32 keyframes become single-item "long" tiles
    prob = synth.make_problem(seed=5, n_kf=60, n_fixed_head=2, n_points=120, mean_track=30.0)
    plan = capi.debug_plan(prob)
    assert plan["tiles"][:, T["is_long"]].sum() > 0
    check_plan(prob, plan, [0, prob.n_pose], [0, prob.n_obs])


def test_plan_big_window_is_reordered_by_first_free_pose():
    prob = synth.config_c3(seed=1, scale=0.1)
    assert (prob.pose_fixed == 0).sum() > 128
    plan = capi.debug_plan(prob)
    assert plan["reordered"] and not plan["pq_shared"] and plan["smallwin"]
    order = plan["landmark_order"]
    assert np.array_equal(np.sort(order), np.arange(prob.n_point))  # a bijection
    check_plan(prob, plan, [0, prob.n_pose], [0, prob.n_obs])
    # internal landmarks are sorted by their first (smallest) free pose slot, ties in the caller's order
    slot = free_slots(prob, [0, prob.n_pose])
    big = np.iinfo(np.int64).max
    s = np.where(slot[prob.obs_pose] >= 0, slot[prob.obs_pose], big)
    first = np.full(prob.n_point, big)
    np.minimum.at(first, prob.obs_point, s)
    key = first[order]
    assert np.all(np.diff(key) >= 0)
    same = np.flatnonzero(np.diff(key) == 0)
    assert np.all(order[same + 1] > order[same])


def test_plan_rejects_bad_input():
    wins = [synth.small_window(seed=s) for s in range(2)]
    prob, pp, tp, op = synth.concat_windows(wins)
    bad = op.copy()
    bad[-1] -= 1  # windows must cover all observations
    with pytest.raises(capi.SqrtBAError):
        capi.debug_plan(prob, pp, tp, bad)
    prob.obs_pose[0] = prob.n_pose  # pose index out of range
    with pytest.raises(capi.SqrtBAError):
        capi.debug_plan(prob, pp, tp, op)
    prob.obs_pose[0] = pp[1]  # an observation of window 0 that points at a pose of window 1
    with pytest.raises(capi.SqrtBAError):
        capi.debug_plan(prob, pp, tp, op)
