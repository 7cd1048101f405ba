set -x
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2m_pytest_full.log
timeout 600 python bench.py > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1
