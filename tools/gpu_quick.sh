#!/bin/bash
# quick iteration: parity tests + kernel-only bench lines for a few variants
# usage: tools/gpu_quick.sh tag "env1|args1" "env2|args2" ...
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
if [ $# -eq 0 ]; then set -- "|" "ADT_FIR_KERNEL=p32|" "|--workload eq" "|--workload eq --fft-size 8192" "ADT_FIR_KERNEL=p32|--workload eq --fft-size 8192"; fi
for spec in "$@"; do
  envs="${spec%%|*}"; args="${spec#*|}"
  env $envs python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e $args 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('[$envs] [$args]', d['config']['fft_size'], d['config']['hop'], 'ms', round(d['ms_per_step'],4), 'Ms/s', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'rms', d['config']['parity_rms_vs_oracle'])
" | tee -a $OUT/bench_variants.txt
done
