# usage: bash tools/power_ab.sh name1 name2 ...   (names of build_ab/libadaprox_<name>.so; "base" = the shipped library)
# per variant: bench value, ms per step, and the GPU's power / SM clock while it ran
mark() { echo "$(date +%H:%M:%S.%N | cut -c1-12) $1" >> /tmp/marks.log; }
nvidia-smi --query-gpu=timestamp,power.draw,clocks.sm,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 20 > /tmp/pw.log &
SM=$!
rm -f /tmp/marks.log; sleep 0.5
for v in "$@"; do
  if [ $v = base ]; then unset ADAPROX_LIB; else export ADAPROX_LIB=/root/repo/build_ab/libadaprox_$v.so; fi
  mark ${v}_start
  python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 3 --power-iters 2 > /tmp/ab_$v.json 2>/tmp/ab_$v.err
  mark ${v}_end
done
kill $SM
python - "$@" <<'PY'
import json, sys
def tsec(s):
    h,m,rest=s.split(':'); return int(h)*3600+int(m)*60+float(rest)
marks=[(tsec(l.split()[0]), l.split()[1]) for l in open('/tmp/marks.log')]
rows=[]
for l in open('/tmp/pw.log'):
    p=[x.strip() for x in l.split(',')]
    try: rows.append((tsec(p[0].split()[1]), float(p[1]), float(p[2]), p[3].startswith('Active')))
    except Exception: pass
for (t0,n0),(t1,n1) in zip(marks[::2], marks[1::2]):
    v=n0[:-6]
    busy=[r for r in rows if t0<=r[0]<=t1 and r[1]>450]
    try:
        d=json.load(open('/tmp/ab_%s.json' % v)); val='%.1f it/s %.2f ms' % (d['value'], d['ms_per_step'])
    except Exception as e:
        val='FAILED '+open('/tmp/ab_%s.err' % v).read()[-200:]
    if busy:
        print(json.dumps(dict(variant=v, bench=val, power_avg_w=round(sum(r[1] for r in busy)/len(busy)), power_max_w=round(max(r[1] for r in busy)),
              sm_mhz_avg=round(sum(r[2] for r in busy)/len(busy)), capped_pct=round(100*sum(r[3] for r in busy)/len(busy)))))
    else: print(v, val, 'no busy samples')
PY
