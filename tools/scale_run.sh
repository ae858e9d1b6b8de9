# one 8-GPU box: bench.py at N = 8 / 4 / 2 (row-sharded headline instance), row-sharded LAD (AdaPDM+) and the lambda path at N = 8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "n8 exit $?"
$TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 40 --warmup 5 --to-tol 0 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "n4 exit $?"
$TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 40 --warmup 5 --to-tol 0 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 exit $?"
$TR --nproc-per-node 8 --master-port 29544 tools/bench_pd_multi.py > gpurun_out/r02_c3_lad_sharded_n8.jsonl 2> gpurun_out/r02_c3_lad_n8.err; echo "lad8 exit $?"
$TR --nproc-per-node 2 --master-port 29545 tools/bench_pd_multi.py > gpurun_out/r02_c3_lad_sharded_n2.jsonl 2> gpurun_out/r02_c3_lad_n2.err; echo "lad2 exit $?"
$TR --nproc-per-node 8 --master-port 29546 tools/bench_path_multi.py > gpurun_out/r02_c5_lambda_path_n8.json 2> gpurun_out/r02_c5_n8.err; echo "path8 exit $?"
grep -h "^{" gpurun_out/r02_c3_lad_sharded_n8.jsonl gpurun_out/r02_c3_lad_sharded_n2.jsonl gpurun_out/r02_c5_lambda_path_n8.json | cut -c1-400
