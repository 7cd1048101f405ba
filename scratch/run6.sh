set -x
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision f16 --no-cpu --no-other > gpurun_out/r2f_bench_f16.json 2> gpurun_out/r2f_bench_f16.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
