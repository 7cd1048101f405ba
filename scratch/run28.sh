set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "edge_shapes" 2>&1 | grep -v "^$" | tail -6 > gpurun_out/s17_pytest_k.log
bash scratch/run17.sh s17
timeout 200 python bench.py --config 5 --precision f16 --no-cpu --no-other > gpurun_out/s17_bench_c5_f16.json 2> gpurun_out/s17_bench_c5_f16.err
