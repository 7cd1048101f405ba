set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_config1 and not full_pair_list" 2>&1 | tail -15 > gpurun_out/r2c_pytest.log
for p in f16 f16x3 tf32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $p --no-cpu --no-other > gpurun_out/r2c_bench_$p.json 2> gpurun_out/r2c_bench_$p.err
done
for nw in 5 7 8; do
  VLG_TC_WINDOW=$nw timeout 300 python bench.py --steps 10 --warmup 3 --precision f16 --no-cpu --no-other > gpurun_out/r2c_bench_f16_w$nw.json 2> gpurun_out/r2c_bench_f16_w$nw.err
done
