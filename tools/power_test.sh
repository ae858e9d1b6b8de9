# power / clock of the GPU while (a) the HBM->shared stream probe and (b) the bench solve run (helpers off / on)
nvidia-smi --query-gpu=timestamp,power.draw,clocks.sm,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 20 > /tmp/pw.log &
SM=$!
mark() { echo "$(date +%H:%M:%S.%N | cut -c1-12) $1" >> /tmp/marks.log; }
rm -f /tmp/marks.log; sleep 0.5
mark probe_start; timeout 100 tools/probes/ingest_probe > /tmp/ingest.log 2>&1; mark probe_end
sleep 1
mark bench0_start; ADAPROX_HELPERS=0 python bench.py --steps 60 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 3 --power-iters 2 > /tmp/b0.json 2>/dev/null; mark bench0_end
kill $SM
python - <<'PY'
import json
def tsec(s):
    h,m,rest=s.split(':'); return int(h)*3600+int(m)*60+float(rest)
marks=[(tsec(l.split()[0]), l.split()[1]) for l in open('/tmp/marks.log')]
rows=[]
for l in open('/tmp/pw.log'):
    p=[x.strip() for x in l.split(',')]
    try: rows.append((tsec(p[0].split()[1]), float(p[1]), float(p[2]), p[3].startswith('Active')))
    except Exception: pass
for (t0,n0),(t1,n1) in zip(marks[::2], marks[1::2]):
    seg=[r for r in rows if t0<=r[0]<=t1]
    busy=[r for r in seg if r[1]>450]
    if busy:
        print(n0, '->', n1, 'busy samples', len(busy), 'power avg %.0f W max %.0f W' % (sum(r[1] for r in busy)/len(busy), max(r[1] for r in busy)),
              'SM clock avg %.0f MHz min %.0f' % (sum(r[2] for r in busy)/len(busy), min(r[2] for r in busy)), 'sw_power_cap in %.0f%% of them' % (100*sum(r[3] for r in busy)/len(busy)))
d=json.load(open('/tmp/b0.json')); print('bench', d['value'], d['clocks'])
PY
