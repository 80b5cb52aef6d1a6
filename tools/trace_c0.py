import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from bench import load_pkg
pkg = load_pkg()
np.set_printoptions(linewidth=200, precision=4, suppress=True)
for rtol in (1e-9, 1e-8, 1e-7):
    ba = pkg.SqrtBA(pcg_rtol=rtol)
    prob = pkg.synth.config_c0(0)
    ba.set_problem(prob)
    st = ba.solve_local()
    tr = ba.trace()
    print("rtol", rtol, "cg iters per trial", tr[:, 8].astype(int).tolist(), "total", int(tr[:,8].sum()), "ms", round(st["ms_total"],2), "launches", st["kernel_launches"])
    print("   final chi2 %.8f" % tr[-1, 5])
    ba.close()
