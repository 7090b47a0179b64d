#!/bin/bash
# quick iteration: parity tests + kernel-only bench lines for a few variants
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for args in "" "OCC3" "--workload eq" "--workload eq --fft-size 8192" "OCC3 --workload eq --fft-size 8192"; do
  if [[ "$args" == OCC3* ]]; then export ADT_FIR_OCC=3; args="${args#OCC3}"; tagx=occ3; else unset ADT_FIR_OCC; tagx=occ2; fi
  python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e $args 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('$tagx $args', d['config']['fft_size'], d['config']['hop'], 'ms', round(d['ms_per_step'],4), 'Ms/s', round(d['value']), 'frac', round(d['roofline']['frac'],4), d['clocks'])
" | tee -a $OUT/bench_variants.txt
done
