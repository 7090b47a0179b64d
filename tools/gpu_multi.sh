#!/bin/bash
# multi-GPU evidence: tools/gpu_multi.sh N tag [extra bench workloads...]
N=${1:-2}; TAG=${2:-multi}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; lscpu | grep -E "NUMA|Model name|Socket|^CPU\(s\)" >> $OUT/topo.txt
cat /sys/fs/cgroup/cpu.max >> $OUT/topo.txt 2>&1
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
ADT_TEST_WORLD=$N NCCL_DEBUG_FILE=$OUT/nccl_test_%p.log python -m pytest tests/test_sharding.py -x -q -m gpu -s > $OUT/nccl_${N}gpu_test.log 2>&1; echo "nccl test rc=$?"; tail -4 $OUT/nccl_${N}gpu_test.log
run() {  # name, gpus, args
  local name=$1 g=$2; shift 2
  NCCL_DEBUG_FILE=$OUT/nccl_${name}_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $g --steps 20 --warmup 3 "$@" > $OUT/bench_${name}.json 2> $OUT/bench_${name}.err; echo "$name rc=$?"
  python - $OUT/bench_${name}.json <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("FAILED", e); sys.exit(0)
print(" value", round(d["value"]), "ms/pass", round(d["plan"]["ms_per_pass"], 4), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
      "e2e ms", round(d["e2e"]["ms_per_step"], 1) if d.get("e2e") else None, "numa", (d.get("e2e") or {}).get("numa"))
sg = d.get("scatter_gather")
if sg: print(" scatter_gather", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sg.items() if k not in ("impl", "note")})
for s in d.get("secondary") or []:
    print("    ", s.get("workload"), s.get("fft_size"), "ms", round(s.get("ms_per_pass", 0), 3), "frac", round(s.get("frac", 0), 4), s.get("error") or "")
PY
}
run lowcut_${N}gpu $N
for w in "$@"; do
  g=${w%%:*}; wl=${w#*:}
  run ${wl}_${g}gpu $g --workload $wl --no-secondary
done
grep -h "NCCL INFO.*\(comm 0x\|Init COMPLETE\|NVLS\|nranks\)" $OUT/nccl_*.log 2>/dev/null | sed 's/^.*NCCL INFO/NCCL INFO/' | sort | uniq -c | sort -rn | head -20 > $OUT/nccl_summary.txt
head -12 $OUT/nccl_summary.txt
