#!/bin/bash
# usage: tools/gpu_multi2.sh N tag [bench args...]   one torchrun bench at N GPUs, compact summary
N=$1; TAG=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 "$@" > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "rc=$?"
python - $OUT/bench_${N}gpu.json <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print("value", round(d["value"]), "ms/pass", round(d["plan"]["ms_per_pass"], 4), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"])
print("e2e", round(e.get("value", 0)), "ms", round(e.get("ms_per_step", 0), 1), "transfer_only", e.get("transfer_only"))
sg = d.get("scatter_gather")
if sg: print("scatter_gather", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sg.items() if k not in ("impl", "note")})
PY
