"""sqrtba -- B200-native square-root Levenberg-Marquardt bundle adjustment (hot path of lutao98/SqrtLM-SLAM).

The product is the C-ABI shared library built from csrc/ (include/sqrtba.h -> libsqrtba.so).  This Python
package is only the thin ctypes mirror used by tests and bench.py, the in-tree build script and the
synthetic-workload generator.  There is no CPU implementation of the solver in here.
"""
from . import synth  # noqa: F401
from . import build  # noqa: F401
from . import capi  # noqa: F401
from .capi import SqrtBA, SqrtBAError  # noqa: F401
from . import host_harness  # noqa: F401
from . import multi  # noqa: F401
