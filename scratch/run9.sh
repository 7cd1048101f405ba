set -x
for p in f16 f16x3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $p --no-cpu --no-other > gpurun_out/r2i_bench_$p.json 2> gpurun_out/r2i_bench_$p.err
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multi_curve or config5 or deterministic or split_row" 2>&1 | tail -5 > gpurun_out/r2i_pytest.log
