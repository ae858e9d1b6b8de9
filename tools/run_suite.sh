python -m pytest tests -q -m gpu 2>&1 | tail -6
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_path_gemm -c 1 -o /tmp/ncu_path python tools/profile_kernels.py path > gpurun_out/ncu_k_path_gemm.log 2>&1
python tools/ncu_summary.py /tmp/ncu_path.ncu-rep gpurun_out/r02_ncu_k_path_gemm_AX.json 1 "ncu --set full --clock-control none: python tools/profile_kernels.py path (first launch of k_path_gemm<1,4>: one K slab of R = A X - b, 16384 x 8192 x 256, BK = 32, 3-stage ring)" > /dev/null 2>> gpurun_out/ncu_k_path_gemm.log || echo "summary failed"
tail -n 5 gpurun_out/ncu_k_path_gemm.log
ls -la gpurun_out/
