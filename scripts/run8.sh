set -x
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/h2d_ceiling.py 2>&1 | grep -v "^W\|^\*" | tail -$((n+1))
done > gpurun_out/r02_h2d_ceiling.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -3 gpurun_out/r02_bench_8gpu.err
cat gpurun_out/r02_h2d_ceiling.txt | tail -20
cut -c1-400 gpurun_out/r02_bench_8gpu.json
