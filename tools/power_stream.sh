mark() { echo "$(date +%H:%M:%S.%N | cut -c1-12) $1" >> /tmp/marks.log; }
nvidia-smi --query-gpu=timestamp,power.draw,clocks.sm,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 20 > /tmp/pw.log &
SM=$!
rm -f /tmp/marks.log; sleep 0.5
for cfg in "112 64 3 0" "112 64 3 1" "112 64 3 2" "148 64 3 1" "148 64 3 2"; do
  set -- $cfg
  mark "s$1_$4_start"; tools/probes/ingest_probe steady $1 $2 $3 $4 3 >> /tmp/steady.log; mark "s$1_$4_end"; sleep 1
done
kill $SM
cat /tmp/steady.log
python - <<'PY'
def tsec(s):
    h,m,rest=s.split(':'); return int(h)*3600+int(m)*60+float(rest)
marks=[(tsec(l.split()[0]), l.split()[1]) for l in open('/tmp/marks.log')]
rows=[]
for l in open('/tmp/pw.log'):
    p=[x.strip() for x in l.split(',')]
    try: rows.append((tsec(p[0].split()[1]), float(p[1]), float(p[2]), p[3].startswith('Active')))
    except Exception: pass
for (t0,n0),(t1,n1) in zip(marks[::2], marks[1::2]):
    seg=[r for r in rows if t0+1.0<=r[0]<=t1-0.2]      # steady part
    if seg: print(n0, 'samples', len(seg), 'power avg %.0f W max %.0f' % (sum(r[1] for r in seg)/len(seg), max(r[1] for r in seg)), 'SM clock avg %.0f' % (sum(r[2] for r in seg)/len(seg)), 'capped %.0f%%' % (100*sum(r[3] for r in seg)/len(seg)))
PY
