#!/bin/bash
# round-2 final GPU call: the whole GPU suite, then the bench line the driver will ask for
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2f_gputest.log
tail -3 gpurun_out/r2f_gputest.log
timeout 420 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
echo "bench rc $?"; tail -c 400 gpurun_out/r2f_bench_n1.err
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'single', d['single_step']['value'], 'e2e', d['e2e']['value'], d['e2e']['steps_in_flight'], 'frac', d['roofline']['frac'],
          'parity', d['parity_check'], 'cpu', d['cpu_baseline']['value'], {k: v['value'] for k, v in (d['extra_configs'] or {}).items()}, 'wall', d['setup_s'])
except Exception as e:
    print('no bench line', e)
P
