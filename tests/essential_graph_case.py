"""A small map after a loop closure, the way LoopClosing::CorrectLoop hands it to Optimizer::OptimizeEssentialGraph
(src/backend/LoopClosing.cc:863), plus the graph the reference's rules make of it (g2oOptimizer.cc:1258-1448), stated
independently of the adapter.  Shared by the CPU gather test and the GPU adapter test."""
import numpy as np


def build(pkg, synth, n_kf=40, seed=3, fix_scale=True):
    est, fixed, _, _ = synth.pose_graph(seed, n_kf=n_kf, fix_scale=fix_scale, n_loop=0)
    truth, _, _, _ = synth.pose_graph(seed, n_kf=n_kf, fix_scale=fix_scale, n_loop=0)
    rng = np.random.default_rng(seed)
    T = np.tile(np.eye(4, dtype=np.float32), (n_kf, 1, 1))
    for k in range(n_kf):
        T[k, :3, :3] = synth.quat_to_rotmat(est[k, :4]).astype(np.float32)
        T[k, :3, 3] = est[k, 4:7].astype(np.float32)
    # map points: a few per keyframe, in front of it; reference keyframe = that keyframe
    n_per = 3
    pts, ref = [], []
    for k in range(n_kf):
        Rcw, tcw = T[k, :3, :3].astype(np.float64), T[k, :3, 3].astype(np.float64)
        for _ in range(n_per):
            Xc = np.array([rng.uniform(-3, 3), rng.uniform(-1, 1), rng.uniform(5, 20)])
            pts.append(Rcw.T @ (Xc - tcw))
            ref.append(k)
    pts = np.array(pts, np.float32)
    m = pkg.host_harness.MockMap.from_poses(T, pts)
    loop_kf, cur_kf = 0, n_kf - 1
    weights = {}
    for k in range(1, n_kf):
        m.set_parent(k, k - 1)
        weights[(k, k - 1)] = 200
    for k in range(2, n_kf):       # covisibility: strong to k-2 for even k, weak (no edge) to k-3
        weights[(k, k - 2)] = 150 if k % 2 == 0 else 60
        if k >= 3:
            weights[(k, k - 3)] = 99
    m.add_loop_edge(12, 3)         # a loop closed earlier
    weights[(12, 3)] = 180
    # the loop found now: current keyframe + two neighbours were corrected by Sim3 propagation
    corrected, non_corrected = {}, {}
    for k in (cur_kf, cur_kf - 1, cur_kf - 2):
        non_corrected[k] = est[k].copy()
        c = truth[k].copy()
        c[4:7] += rng.normal(0, 0.01, 3)
        c[7] = 1.0 if fix_scale else 1.02
        corrected[k] = c
    connections = [(cur_kf, loop_kf), (cur_kf, 1), (cur_kf - 1, 1), (cur_kf - 2, 2)]
    weights[(cur_kf, loop_kf)] = 10     # exempt from the weight test (the loop pair itself)
    weights[(cur_kf, 1)] = 120
    weights[(cur_kf - 1, 1)] = 130
    weights[(cur_kf - 2, 2)] = 40        # too weak: no edge
    for (a, b), w in weights.items():
        m.set_weight(a, b, w)
    for j, k in enumerate(ref):
        if k == cur_kf - 1 and j % 2 == 0:   # corrected by the loop fusion: reference taken from mnCorrectedReference
            m.set_ref_kf(j, k, corrected_by=cur_kf, corrected_ref=cur_kf)
            ref[j] = cur_kf
        else:
            m.set_ref_kf(j, k)
    return dict(map=m, T=T, pts=pts, ref=np.array(ref), loop_kf=loop_kf, cur_kf=cur_kf, corrected=corrected,
                non_corrected=non_corrected, connections=connections, weights=weights, n_kf=n_kf, fix_scale=fix_scale)


def expected_graph(case, synth):
    """The reference's rules restated on the Python side: returns (vert8, edges [(i, j)], meas8 list)."""
    n, T = case["n_kf"], case["T"]
    cur, loop = case["cur_kf"], case["loop_kf"]
    W = dict(case["weights"])
    W.update({(b, a): w for (a, b), w in case["weights"].items()})
    vert = np.zeros((n, 8))
    for k in range(n):
        if k in case["corrected"]:
            vert[k] = case["corrected"][k]
        else:
            q = synth.rotmat_to_quat_eigen(T[k, :3, :3].astype(np.float64)[None], normalize=False)[0]
            vert[k] = np.concatenate([q, T[k, :3, 3].astype(np.float64), [1.0]])
    edges, meas, inserted = [], [], set()
    conn = {}
    for a, b in case["connections"]:
        conn.setdefault(a, set()).add(b)
    for i in sorted(conn):                                      # std::map<KeyFrame*, ...>: pointer order = creation order here
        Swi = synth.sim3_inv(vert[i])
        for j in sorted(conn[i]):
            if (i != cur or j != loop) and W.get((i, j), 0) < 100:
                continue
            edges.append((i, j)); meas.append(synth.sim3_mul(vert[j], Swi))
            inserted.add((min(i, j), max(i, j)))
    prior = lambda k: case["non_corrected"].get(k, vert[k])
    loops = {12: {3}, 3: {12}}
    for i in range(n):
        Swi = synth.sim3_inv(prior(i))
        if i >= 1:
            edges.append((i, i - 1)); meas.append(synth.sim3_mul(prior(i - 1), Swi))
        for l in sorted(loops.get(i, ())):
            if l < i:
                edges.append((i, l)); meas.append(synth.sim3_mul(prior(l), Swi))
        strong = sorted([(-w, j) for (a, j), w in W.items() if a == i and w >= 100])
        for _, j in strong:
            if j == i - 1 or j == i + 1 or j in loops.get(i, ()):   # parent, child, loop edge
                continue
            if j < i and (min(i, j), max(i, j)) not in inserted:
                edges.append((i, j)); meas.append(synth.sim3_mul(prior(j), Swi))
    return vert, edges, meas
