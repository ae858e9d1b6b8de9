# how many MB of a dense sweep's tail are loaded evict_last for the next (reversed) sweep: ADAPROX_L2_KEEP_MB (0 = every tile evict_first, shipped; -1 = no eviction hints)
for mb in ${KEEPS:-0 24 48 64 96}; do
  echo "== ADAPROX_L2_KEEP_MB=$mb"
  ADAPROX_L2_KEEP_MB=$mb python tools/bench_configs.py ${CONFIGS:-lad svmgram} 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], '| it', d['iterations'], '| us/it', round(d['us_per_iteration'],1), '| GB/s', round(d.get('hbm_gbs',0)), '| res', d['final_norm_res'])
    else: print(l.rstrip())"
done
