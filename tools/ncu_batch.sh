# ncu captures of the hot kernels on one B200; summaries (JSON) go to gpurun_out/, the .ncu-rep files are deleted on the box
# (they exceed the 64 MiB that travel back).  Usage on the GPU box:  bash tools/ncu_batch.sh
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, launch-skip, grad evals, description, command...
  local name=$1 kre=$2 skip=$3 evals=$4 what=$5; shift 5
  $NCU -k regex:$kre -s $skip -c 1 -o /tmp/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu_$name.ncu-rep gpurun_out/r02_ncu_$name.json $evals "$what" > /dev/null 2>> gpurun_out/ncu_$name.log || echo "summary failed: $name"
  rm -f /tmp/ncu_$name.ncu-rep
}
cap k_adapgm_resident k_adapgm_resident 0 2001 "ncu --set full --clock-control none: python tools/profile_kernels.py resident (configs[0] lasso 400x1000, 2000 iterations in one launch)" python tools/profile_kernels.py resident
cap k_primal_dual_csr k_primal_dual 0 201 "ncu --set full --clock-control none: python tools/profile_kernels.py c2 (configs[1] sparse logreg 20242x47236 CSR, 200 iterations: the spmv_rows sweeps inside k_primal_dual<false>)" python tools/profile_kernels.py c2
cap k_proxgrad_family k_proxgrad_family 0 300 "ncu --set full --clock-control none: python tools/profile_kernels.py pgfamily (backtracking PG, lasso 4000x1000, 300 iterations)" python tools/profile_kernels.py pgfamily
cap k_malitsky_pock k_malitsky_pock 0 60 "ncu --set full --clock-control none: python tools/profile_kernels.py mp (Malitsky-Pock on LAD 50000x2001, 60 iterations)" python tools/profile_kernels.py mp
cap k_path_gemm_AX "k_path_gemm.*1.*4" 0 1 "ncu --set full --clock-control none: python tools/profile_kernels.py path (first launch of k_path_gemm<1,4>: R = A X - b, 16384 x 8192 x 256, BK = 32)" python tools/profile_kernels.py path
ADAPROX_FUSED_NONCOOP=1 cap k_adapgm_fused k_adapgm_fused 1 4 "ncu --set full --clock-control none, ADAPROX_FUSED_NONCOOP=1 (ncu cannot replay cooperative + cluster launches): bench.py --steps 3 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 1 --power-iters 2; the captured launch is the 3-iteration timed solve = 4 gradient evaluations" python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 1 --power-iters 2
ADAPROX_FUSED_NONCOOP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ncu.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-configs --to-tol 0 --reps 1 --power-iters 2 > gpurun_out/ncu_launches.log 2>&1
python tools/phase_timing.py > gpurun_out/r02_phase_timing.log 2>&1
ls -la gpurun_out/
