set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2s_pytest_full.log
timeout 400 python bench.py > gpurun_out/r2s_bench_default.json 2> gpurun_out/r2s_bench_default.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
timeout 200 python bench.py --config 5 --no-cpu --no-other > gpurun_out/r2s_bench_c5.json 2> gpurun_out/r2s_bench_c5.err
