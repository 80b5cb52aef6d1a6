#!/bin/bash
for n in 1 2 3; do echo "persist CTAs/SM $n"; for c in c0 c2; do SQRTBA_PERSIST_CTAS_PER_SM=$n python tools/single_window.py --config $c | cut -c1-170; done; done
for n in 2 3; do echo "proxy nshards 8, persist CTAs/SM $n"; SQRTBA_PERSIST_CTAS_PER_SM=$n python tools/gba_proxy.py --nshards 8 2>&1 | tail -1 | cut -c150-420; done
