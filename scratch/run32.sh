set -x
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not config3_final and not full_config1 and not full_pair_list" 2>&1 | tail -4 > gpurun_out/s21_pytest.log
timeout 200 python bench.py --config 5 --precision f16 --no-cpu --no-other > gpurun_out/s21_bench_c5_f16.json 2> gpurun_out/s21_bench_c5_f16.err
timeout 200 python bench.py --config 5 --precision f16x3 --no-cpu --no-other > gpurun_out/s21_bench_c5_f16x3.json 2> gpurun_out/s21_bench_c5_f16x3.err
