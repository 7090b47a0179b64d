#!/bin/bash
# round-2 call A (1 GPU): parity tests, sanitizer, bench with secondary list
OUT=gpurun_out/r2a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; cat /sys/fs/cgroup/cpu.max >> $OUT/host.txt 2>&1; lscpu | head -25 >> $OUT/host.txt; numactl -H >> $OUT/host.txt 2>&1
python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
ADT_LIB_PATH=$PWD/pyaudiodsptools_b200/libadt_b200_ab.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "every_kernel_variant" > $OUT/pytest_gpu_ab.log 2>&1; echo "pytest ab rc=$?"; tail -2 $OUT/pytest_gpu_ab.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2a/bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],4),'ms_pass',round(d['plan']['ms_per_pass'],4),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value'],1),d['cpu_baseline']['kind'],d['cpu_baseline']['environment'],'single',round(d['cpu_baseline']['single_core']['value'],2))
print('clocks',d['clocks'])
for s in d['secondary']: print(s.get('workload'), s.get('fft_size'), s.get('hop'), round(s.get('ms_per_pass',0),3), round(s.get('frac',0),4), s.get('parity_rms_vs_oracle'), s.get('error'))
PY
python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $OUT/bench_ref.json
tools/sanitize.sh r2a
