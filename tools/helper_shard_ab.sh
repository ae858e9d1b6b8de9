# emulate the per-rank sweep of the sharded runs on one GPU: m = 8192 (N = 8 shard), 16384 (N = 4), 65536 (N = 1); helpers off / on / hold variants
for m in 8192 16384 65536; do
  for cfg in "0 13" "2 13" "2 0" "2 20"; do
    set -- $cfg
    ADAPROX_HELPERS=$1 ADAPROX_HELPER_HOLD=$2 python bench.py --m $m --steps 40 --warmup 5 --no-cpu --no-configs --to-tol 0 --power-iters 2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('m=$m helpers=$1 hold=$2', round(d['value'],1), 'it/s', round(d['ms_per_step'],3), 'ms', [round(x,3) for x in d['repetitions']['ms_per_step']], d['clocks']['sm_mhz'])"
  done
done
