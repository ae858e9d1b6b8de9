python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref exit $?"
NCU="ncu --set full --clock-control none --import-source on -f"
ADAPROX_FUSED_NONCOOP=1 ADAPROX_HELPERS=0 $NCU -k regex:k_adapgm_fused -s 1 -c 1 -o /tmp/ncu_fused python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 1 --power-iters 2 > gpurun_out/ncu_k_adapgm_fused.log 2>&1
python tools/ncu_summary.py /tmp/ncu_fused.ncu-rep gpurun_out/r02_ncu_k_adapgm_fused.json 4 "ncu --set full --clock-control none, ADAPROX_FUSED_NONCOOP=1 ADAPROX_HELPERS=0 (ncu cannot replay cooperative + cluster launches and serialises kernels, so the helper CTAs cannot run beside it): bench.py --steps 3 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 1 --power-iters 2; the captured launch is the 3-iteration timed solve = 4 gradient evaluations; final kernel of round 2 (one reducer warp, suspended pollers, L2 evict_first)" > /dev/null 2>> gpurun_out/ncu_k_adapgm_fused.log || echo "summary failed"
ADAPROX_FUSED_NONCOOP=1 ADAPROX_HELPERS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ncu.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 1 --power-iters 2 > gpurun_out/ncu_launches.log 2>&1
python tools/phase_timing.py > gpurun_out/r02_phase_timing.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["repetitions"]["ms_per_step"], d["roofline"]["frac"], d["time_to_tol"]["seconds"], d["cpu_baseline"]["value"], d["clocks"])
for k,v in d["configs"].items(): print(k, {a:b for a,b in v.items() if a in ("us_per_iteration","ms_per_batched_iteration")}, v.get("roofline",{}).get("frac"))
r=json.load(open("gpurun_out/r02_bench_reference_arm.json")); print("reference arm", r["value"], r["ms_per_step"], r["extrapolated"], r["timed_region_s"], r["cpu_baseline"]["cores"])
PY
