set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2r_bench_n4.json 2> gpurun_out/r2r_bench_n4.err
