#!/bin/bash
# usage: tools/gpu_variants.sh tag "ENV=.. ENV2=..|bench args" ...   -> one compact line per variant
TAG=${1:-v}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
i=0
for spec in "$@"; do
  envs="${spec%%|*}"; args="${spec#*|}"; i=$((i+1))
  env $envs python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e $args > $OUT/v$i.json 2> $OUT/v$i.err || tail -3 $OUT/v$i.err
  python - "$OUT/v$i.json" "[$envs] [$args]" <<'PY' | tee -a $OUT/variants.txt
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[2], "FAILED", e); sys.exit(0)
c = d.get("clocks") or {}
print(sys.argv[2], "N", d["plan"]["fft_size"], "hop", d["plan"]["hop"], "ms/pass", round(d["plan"]["ms_per_pass"], 4),
      "frac", round(d["roofline"]["frac"], 4), "rms", "%.1e" % d["plan"]["parity_rms_vs_oracle"], "sm_mhz", c.get("sm_mhz"), c.get("reasons"))
for s in d.get("secondary") or []:
    print("    ", s.get("workload"), s.get("fft_size"), s.get("hop"), "ms", round(s.get("ms_per_pass", 0), 3), "frac", round(s.get("frac", 0), 4),
          s.get("error") or "", s.get("bit_exact_vs_oracle_first_20000", ""), round(s.get("Msamples_s_per_band", 0)))
PY
done
