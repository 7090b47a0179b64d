for mb in 8 24 48 128 512; do
ADT_FIR_GROUP_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('group_mb', $mb, 'e2e ms', round(d['e2e']['ms_per_step'],2), 'Msamples/s', round(d['e2e']['value']))"
done
