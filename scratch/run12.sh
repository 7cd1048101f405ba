set -x
for nw in 8 9 10 12; do
  VLG_TC_WINDOW=$nw timeout 300 python bench.py --steps 10 --warmup 3 --precision f16x3 --no-cpu --no-other > gpurun_out/r2l_bench_x3_w$nw.json 2> gpurun_out/r2l_bench_x3_w$nw.err
done
timeout 600 python -m pytest tests/test_gpu_dropin.py -q -m gpu -x -k "single_decoder" 2>&1 | tail -3 > gpurun_out/r2l_pytest_single.log
