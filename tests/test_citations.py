"""Every `file:line` citation of the reference in the docs, headers, kernels, adapter and oracle must name a file that
exists in the reference checkout and a line range inside it (runs where /root/reference is present)."""
import collections
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OWN = {"sqrtba.h", "sqrtbaOptimizer.cc", "Optimizer.h", "refba.cpp", "map_types.h", "sqrtba_solver.cu", "host_pool.h"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not present on this machine")
def test_reference_citations_resolve():
    index = collections.defaultdict(list)
    for dp, _, fs in os.walk(REF):
        if ".git" in dp:
            continue
        for f in fs:
            index[f].append(os.path.join(dp, f))
    pat = re.compile(r"([A-Za-z0-9_\-/]+\.(?:cc|cpp|h|hpp|yaml|txt))[:`]*:(\d+)(?:-(\d+))?")
    files = [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md", "README.md", "include/sqrtba.h")]
    for sub in ("oracle/*.cpp", "oracle/*.py", "sqrtlm-slam_b200/csrc/*", "sqrtlm-slam_b200/host/*", "sqrtlm-slam_b200/*.py",
                "tests/*.py"):
        files += glob.glob(os.path.join(ROOT, sub))
    lens, bad, n = {}, [], 0
    for fn in files:
        for m in pat.finditer(open(fn, errors="ignore").read()):
            name, a, b = os.path.basename(m.group(1)), int(m.group(2)), int(m.group(3) or m.group(2))
            if name in OWN:
                continue
            cands = index.get(name)
            if not cands:
                bad.append((os.path.relpath(fn, ROOT), m.group(0), "no such file in the reference"))
                continue
            n += 1
            for c in cands:
                if c not in lens:
                    lens[c] = sum(1 for _ in open(c, errors="ignore"))
            if not any(a <= b <= lens[c] for c in cands):
                bad.append((os.path.relpath(fn, ROOT), m.group(0), f"file has {[lens[c] for c in cands]} lines"))
    assert n > 200, n
    assert not bad, bad
