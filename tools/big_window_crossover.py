"""Where does the big-window PCG path (chunked vector phase, two-level preconditioner) start to pay for ONE local window?
C2-shaped windows (one fixed keyframe, ~500 points and ~6000 observations per keyframe) of growing size, both paths."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import sys, json
sys.path.insert(0, %r)
from bench import load_pkg
pkg = load_pkg()
n_kf = int(sys.argv[1])
prob = pkg.synth.make_problem(3, n_kf, 1, 500 * n_kf, 12.0, stereo=True, cand_halfwidth=40)
ba = pkg.SqrtBA()
ba.set_problem(prob)
best = 1e9
for _ in range(4):
    ba.reset_state()
    st = ba.solve_local()
    best = min(best, st["ms_total"])
print(json.dumps(dict(n_free=prob.n_free, n_obs=prob.n_obs, ms=best, cg=st["cg_iters_total"], chunk=st["chunk_precond"], trials=st["lm_trials"])))
""" % ROOT
for n_kf in (41, 61, 81, 100, 121):
    row = {}
    for thr in (129, 2):
        env = dict(os.environ, SQRTBA_BIG_MIN_SLOTS=str(thr))
        out = subprocess.run([sys.executable, "-c", CODE, str(n_kf)], env=env, capture_output=True, text=True).stdout
        row["small_path" if thr == 129 else "big_path"] = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    print(json.dumps(row), flush=True)
