# chunk granularity of the dynamic schedule on a shard-sized sweep (m = 8192 rows on one GPU = the N = 8 rank's work)
for c in 1171 586 391 293 196 147; do
  ADAPROX_FUSED_CHUNK=$c python bench.py --m 8192 --steps 40 --warmup 5 --no-cpu --no-configs --to-tol 0 --power-iters 2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunk_rows=$c', round(d['value'],1), 'it/s', [round(x,3) for x in d['repetitions']['ms_per_step']], d['clocks']['sm_mhz'])"
done
