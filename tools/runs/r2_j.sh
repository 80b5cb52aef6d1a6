#!/bin/bash
SQRTBA_LIB=sqrtlm-slam_b200/libsqrtba_prof.so python tools/prof_stage.py 256 5 2>&1 | grep -i "pipe prof\|matvec"
SQRTBA_LIB=sqrtlm-slam_b200/libsqrtba_prof.so SQRTBA_PIPE_STAGES=3 python tools/prof_stage.py 256 5 2>&1 | grep -i "pipe prof\|matvec"
