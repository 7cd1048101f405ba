set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_config1 and not full_pair_list" 2>&1 | tail -5 > gpurun_out/r2e_pytest.log
VLG_TC_XL2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_config1 and not full_pair_list" 2>&1 | tail -5 > gpurun_out/r2e_pytest_xl2.log
for p in f16 f16x3 tf32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $p --no-cpu --no-other > gpurun_out/r2e_bench_$p.json 2> gpurun_out/r2e_bench_$p.err
done
VLG_TC_XL2=1 timeout 300 python bench.py --steps 10 --warmup 3 --precision f16 --no-cpu --no-other > gpurun_out/r2e_bench_f16_xl2.json 2> gpurun_out/r2e_bench_f16_xl2.err
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --precision f16 --no-cpu --no-other > gpurun_out/r2e_bench_c5_f16.json 2> gpurun_out/r2e_bench_c5_f16.err
