# last verification of the round on one GPU: parity tests, smoke, the default bench line (-> gpurun_out/r02_bench_n1_final.json)
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1_final.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["repetitions"]["ms_per_step"], d["roofline"]["frac"], d["time_to_tol"]["seconds"], d["cpu_baseline"]["value"], d["clocks"], d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, {a:b for a,b in v.items() if a in ("us_per_iteration","ms_per_batched_iteration","iterations")}, v.get("roofline",{}).get("frac"), v.get("cpu_port",{}).get("us_per_iteration"))
PY
tail -3 gpurun_out/r02_bench_n1_final.err
