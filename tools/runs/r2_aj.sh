#!/bin/bash
# round 2, step aj: how much PCG tolerance do the result tolerances leave on C0 windows (seeds 0-5)?
set -x
python tools/rtol_probe_local.py 0,1,2,3,4,5 2>&1 | grep "^{"
