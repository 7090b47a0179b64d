#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, the ncu launch list and
# one full ncu capture of the FIR kernel.  Outputs land in gpurun_out/.
# usage: tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
# launch list (every launch with its device time; shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --channels 200 > $OUT/launches_run.log 2>&1
# full capture of the dominant kernel (one launch, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:fir_block -s 3 -c 1 -o $OUT/fir_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_full_run.log 2>&1
ls -la $OUT
