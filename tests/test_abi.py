"""CPU-only checks of the drop-in boundary: the library builds, loads without a GPU, exports every symbol that
include/sqrtba.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sqrtba.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sqrtba_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    pkg.build.build_lib()
    L = C.CDLL(pkg.capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sqrtba.h but not exported"


def test_struct_layouts_match_header(pkg):
    assert C.sizeof(pkg.capi.Config) == 4 + 4 + 8 + 4 * 4 + 8 * 4  # device(+pad), rtol, 4 ints, reserved[8]
    assert C.sizeof(pkg.capi.Stats) == 4 * 4 + 8 * 7 + 8 * 8
    assert len(pkg.capi.TRACE_COLS) == 10


def test_ctypes_structs_match_the_compiled_header(pkg, tmp_path):
    """sizeof / offsetof of every struct of include/sqrtba.h as gcc sees them, against the ctypes mirrors in capi.py
    (the header is plain C: it must also compile as C)."""
    import subprocess
    structs = {"sqrtba_config": pkg.capi.Config, "sqrtba_stats": pkg.capi.Stats, "sqrtba_lidar": pkg.capi.Lidar}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "sqrtba.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, field, val = ln.split()
        ct = structs[cname]
        want = C.sizeof(ct) if field == "size" else getattr(ct, field).offset
        assert int(val) == want, f"{cname}.{field}: header {val}, ctypes {want}"
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())


def test_no_cpu_fallback_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.SqrtBAError):
        pkg.SqrtBA()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    bad = []
    for base in (os.path.join(ROOT, "sqrtlm-slam_b200"), os.path.join(ROOT, "include")):
        for dp, _, fs in os.walk(base):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"\boracle\b|refba", s):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_hot_kernels_carry_the_blackwell_instructions(pkg):
    """SASS of the built library (sm_100a): the matvec / persistent-PCG kernels move their tiles with TMA bulk copies
    (UBLKCP) completed on mbarriers (SYNCS); the pipelined landmark QR and linearisation stage the next tile with
    cp.async (LDGSTS) and prefetch into L2 with the TMA prefetch (UBLKPF).  Guards against a build that silently lost
    them (wrong arch flag, a fallback path)."""
    import collections
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump is not installed")
    pkg.build.build_lib()
    sass = subprocess.run(["cuobjdump", "-sass", pkg.capi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    fn, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        for k in ("UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "DFMA", "REDG.E.ADD.F32x4", "STL", "LDL"):
            if k in line:
                cnt[fn][k] += 1

    def kernels(sub):
        ks = [c for f, c in cnt.items() if sub in f]
        assert ks, f"no kernel named *{sub}* in the library"
        return ks

    for c in kernels("k_matvec_pipe") + kernels("k_pcg_persist"):
        assert c["UBLKCP"] >= 1 and c["SYNCS"] >= 8 and c["DFMA"] > 50, c
    for c in kernels("k_qr_pipe2"):
        assert c["LDGSTS"] >= 14 and c["UBLKPF"] >= 1 and c["DFMA"] > 300, c
    for c in kernels("k_linearize_pipeILb1E"):
        assert c["LDGSTS"] >= 2 and c["UBLKPF"] >= 1 and c["DFMA"] > 150, c
    # global-BA preconditioner: the pair blocks leave the SMs as 16-byte vector reductions; the Gauss-Jordan sweep of a
    # chunk block keeps its 5x6 tile in registers (no local-memory traffic in the sweep: the only stack use is the 6x6
    # fallback, which spills nothing) and the chunk instantiation of the PCG kernel spills nothing either
    for c in kernels("k_chunk_blocks"):
        assert c["REDG.E.ADD.F32x4"] >= 18, c
    for c in kernels("k_pcg_persistILi2ELb1ELb0ELb1E") + kernels("k_pcg_persistILi2ELb1ELb1ELb1E"):
        assert c["STL"] == 0 and c["LDL"] == 0 and c["UBLKCP"] >= 1, c
