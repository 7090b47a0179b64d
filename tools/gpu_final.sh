#!/bin/bash
# final 1-GPU evidence: tests, sanitizer, ncu captures (headline + EQ kernel), launch list, bench
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 tools/sanitize.sh ${1:-final}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_block -s 3 -c 1 -o $OUT/fir_full -f python bench.py --steps 1 --passes 1 --warmup 3 --no-cpu --no-e2e --no-secondary > $OUT/ncu_full_run.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:fir_persist -s 3 -c 1 -o $OUT/eq_full -f python bench.py --workload eq --steps 1 --passes 1 --warmup 3 --no-cpu --no-e2e --no-secondary > $OUT/ncu_eq_run.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --passes 5 --warmup 3 --no-cpu --no-secondary > $OUT/launches_run.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
ls -la $OUT | head -40
