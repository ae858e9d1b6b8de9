# dense ring sweeps: alternating direction (shipped) vs every sweep first row -> last row (-DADAPROX_SWEEP_ONE_WAY), each with the
# stream loaded evict_first (shipped, ADAPROX_L2_KEEP_MB=0) and without eviction hints (-1); same box
for lib in "" build_ab/libadaprox_oneway.so; do
 for mb in 0 -1; do
  echo "== ${lib:-shipped} ADAPROX_L2_KEEP_MB=$mb"
  ADAPROX_L2_KEEP_MB=$mb ADAPROX_LIB=${lib:+$PWD/$lib} python tools/bench_configs.py ${CONFIGS:-lad svm svmgram} 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], '| it', d['iterations'], '| us/it', round(d['us_per_iteration'],1), '| GB/s', round(d.get('hbm_gbs',0)), '| res', d['final_norm_res'])
    else: print(l.rstrip())"
 done
done
