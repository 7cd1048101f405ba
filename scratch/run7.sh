set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_config1 and not full_pair_list and not config3_final" 2>&1 | tail -15 > gpurun_out/r2g_pytest.log
VLG_TC_XL2=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_config1 and not full_pair_list and not config3_final" 2>&1 | tail -15 > gpurun_out/r2g_pytest_xl2.log
for p in f16 f16x3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $p --no-cpu --no-other > gpurun_out/r2g_bench_$p.json 2> gpurun_out/r2g_bench_$p.err
done
for p in f16 f16x3; do
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --precision $p --no-cpu --no-other > gpurun_out/r2g_bench_c5_$p.json 2> gpurun_out/r2g_bench_c5_$p.err
done
for g in 2 4; do
VLG_TC_G=$g timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --precision f16 --no-cpu --no-other > gpurun_out/r2g_bench_c5_f16_g$g.json 2> gpurun_out/r2g_bench_c5_f16_g$g.err
done
