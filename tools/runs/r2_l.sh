#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_l.log 2>&1; tail -12 gpurun_out/r2_pytest_l.log
python tools/stage_roofline.py --config c4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); [print(k, v) for k,v in d['kernels'].items()]"
python tools/single_window.py --config c0; python tools/single_window.py --config c2
python bench.py --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null > gpurun_out/r2_bench_l.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_l.json')); print(d['value'], d['ms_per_step'], d['e2e']); print(json.dumps(d['global_ba'])[:330])"
