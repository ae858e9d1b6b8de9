mkdir -p gpurun_out
python tools/profile_kernels.py gridres > gpurun_out/gridres_plain.log 2>&1; cat gpurun_out/gridres_plain.log | tail -1
ncu --set full --clock-control none --import-source on -f -k regex:k_adapgm_gridres -c 1 -o /tmp/ncu_gridres python tools/profile_kernels.py gridres > gpurun_out/ncu_gridres.log 2>&1
python tools/ncu_summary.py /tmp/ncu_gridres.ncu-rep gpurun_out/r02_ncu_k_adapgm_gridres.json 2001 "ncu --set full --clock-control none: python tools/profile_kernels.py gridres (lasso 4000x1000, the largest instance of lasso/runme.jl:191-195; 2000 iterations in one cooperative launch)" > /dev/null 2>> gpurun_out/ncu_gridres.log || echo "summary failed"
rm -f /tmp/ncu_gridres.ncu-rep
tail -3 gpurun_out/ncu_gridres.log; ls -la gpurun_out/r02_ncu_k_adapgm_gridres.json
