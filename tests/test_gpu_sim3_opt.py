"""OptimizeSim3 on the device (g2oOptimizer.cc:1560-1796) through the C ABI
(sqrtba_optimize_sim3): against the reference's OWN binary (tests/golden/libg2o_vectors.npz: s3o<k>_*, the schedule
oracle/pin_libg2o_graph.py ran with libg2o.so's Levenberg on real EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ
objects) and against the oracle (refba_optimize_sim3, pinned to the same binary) on larger and batched inputs."""
import os

import numpy as np
import pytest

from oracle import refba

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libg2o_vectors.npz")


@pytest.fixture(scope="module")
def ba(pkg):
    h = pkg.SqrtBA()
    yield h
    h.close()


def informative(tr):
    """Leading LM trials of the first pass whose step changes the cost by more than the noise of the reference's own
    numeric Jacobians (delta = 1e-9 central differences on pixel-sized errors: ~1e-6 relative per entry); below that
    accept / reject is rounding noise that the binary itself does not share with a re-ordered sum."""
    rel = np.abs(tr[:, 4] - tr[:, 5]) / tr[0, 4]
    k = 0
    while k < len(tr) and rel[k] > 1e-4:
        k += 1
    return k


def same_sim3(a, b, atol):
    sgn = np.sign(np.sum(a[:4] * b[:4]))
    np.testing.assert_allclose(a[:4] * sgn, b[:4], rtol=0, atol=atol)
    np.testing.assert_allclose(a[4:], b[4:], rtol=0, atol=atol * (1.0 + np.abs(b[4:7]).max()))


@pytest.mark.parametrize("case", [0, 1])
def test_optimize_sim3_matches_reference_binary(ba, case):
    g = np.load(GOLD)
    pre = f"s3o{case}_"
    n = len(g[pre + "p1c"])
    S, keep, n_in, st = ba.optimize_sim3([0, n], g[pre + "s0"], g[pre + "cam8"], g[pre + "p1c"], g[pre + "p2c"],
                                         g[pre + "meas6"], float(g[pre + "th2"]), bool(g[pre + "fix_scale"]))
    assert st["kernel_launches"] == 1
    assert np.array_equal(keep, g[pre + "keep"]) and int(n_in[0]) == int(g[pre + "nIn"])   # the matches the binary keeps
    same_sim3(S[0], g[pre + "s12"], 1e-6)
    tr = ba.optimize_sim3_trace(0)
    lam = g[pre + "lambda"]
    k = informative(tr)
    assert k >= 1
    np.testing.assert_allclose(tr[:k + 1, 3], lam[:k + 1], rtol=1e-6)     # lambda of trial k still follows from trial k-1
    n0 = int(g[pre + "n_trials_pass0"])
    k1 = int(np.argmax(tr[:, 0] == 1))
    np.testing.assert_allclose(tr[k1, 3], lam[n0], rtol=1e-6)             # lambda_0 of the second pass: same matches left
    if bool(g[pre + "fix_scale"]):
        assert S[0, 7] == g[pre + "s0"][7]


@pytest.mark.parametrize("fix_scale", [False, True])
def test_optimize_sim3_batch_matches_oracle(ba, synth, fix_scale):
    """Six candidate pairs of different sizes in one launch (one of them with too few matches)."""
    sizes = [150, 40, 300, 12, 9, 90]
    cases = [synth.sim3_pair(seed=10 + k, n_matches=m, fix_scale=fix_scale, outlier_frac=0.1 + 0.05 * (k % 3))
             for k, m in enumerate(sizes)]
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    S, keep, n_in, st = ba.optimize_sim3(ptr, np.stack([c[0] for c in cases]), np.stack([c[1] for c in cases]),
                                         np.concatenate([c[2] for c in cases]), np.concatenate([c[3] for c in cases]),
                                         np.concatenate([c[4] for c in cases]), 10.0, fix_scale)
    assert st["kernel_launches"] == 1
    for k, c in enumerate(cases):
        So, keep_o, nin_o, tr_o = refba.optimize_sim3(c[0], c[1], c[2], c[3], c[4], 10.0, fix_scale)
        assert int(n_in[k]) == nin_o, k
        assert np.array_equal(keep[ptr[k]:ptr[k + 1]], keep_o), k
        tr = ba.optimize_sim3_trace(k)
        np.testing.assert_allclose(tr[0, 3:6], tr_o[0, 3:6], rtol=1e-7)   # lambda_0, cost before / after the first trial
        if nin_o == 0:                                                    # gave up: S12 comes back untouched
            assert sizes[k] < 12 and np.array_equal(S[k], c[0])
            continue
        same_sim3(S[k], So, 1e-6)
        # the estimate is a useful one: within the noise of the generator around the truth
        same_sim3(S[k], c[5]["S12"], 0.05)
        agree = (keep[ptr[k]:ptr[k + 1]].astype(bool) == ~c[5]["is_outlier"]).mean()
        assert agree > 0.9
    assert (n_in == 0).sum() >= 1 and (n_in > 0).sum() >= 4


def test_optimize_sim3_rejects_bad_arguments(ba, pkg):
    with pytest.raises(pkg.SqrtBAError):
        ba.optimize_sim3([0, 5, 3], np.zeros((2, 8)), np.zeros((2, 8)), np.zeros((5, 3)), np.zeros((5, 3)), np.zeros((5, 6)))
    with pytest.raises(pkg.SqrtBAError):
        ba.optimize_sim3([0, 5], np.zeros((1, 8)), np.zeros((1, 8)), np.zeros((5, 3)), np.zeros((5, 3)), np.zeros((5, 6)), th2=0.0)
