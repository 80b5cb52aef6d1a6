"""sqrtba -- B200-native square-root Levenberg-Marquardt bundle adjustment (hot path of lutao98/SqrtLM-SLAM).

The product is the C-ABI shared library built from csrc/ (include/sqrtba.h).  This Python package is
only the thin ctypes mirror used by tests and bench.py plus the synthetic-workload generator.
"""
from . import synth  # noqa: F401
