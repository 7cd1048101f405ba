set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not config3_final and not full_config1 and not full_pair_list" 2>&1 | tail -4 > gpurun_out/s30_pytest.log
for m in f16x3f f16; do
timeout 90 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/s30_bench_$m.json 2> gpurun_out/s30_bench_$m.err
done
