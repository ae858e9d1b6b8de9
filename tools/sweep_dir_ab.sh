# dense ring sweeps: A*x against the direction of the previous sweep (shipped) vs every sweep first row -> last row
# (ADAPROX_SWEEP_ONE_WAY=1), each with the stream loaded evict_first (shipped, ADAPROX_L2_KEEP_MB=0) and without eviction hints (-1); same box
for ow in "" 1; do
 for mb in 0 -1; do
  if [ -n "$ow" ]; then label="one way"; else label="alternating"; fi
  echo "== $label ADAPROX_L2_KEEP_MB=$mb"
  env ${ow:+ADAPROX_SWEEP_ONE_WAY=1} ADAPROX_L2_KEEP_MB=$mb python tools/bench_configs.py ${CONFIGS:-lad svm svmgram} 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], '| it', d['iterations'], '| us/it', round(d['us_per_iteration'],1), '| GB/s', round(d.get('hbm_gbs',0)), '| res', d['final_norm_res'])
    else: print(l.rstrip())"
 done
done
