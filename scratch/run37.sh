set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "f16x3f and not full_pair_list" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/s28_pytest_mix.log
for m in f16x3f f16x3; do
timeout 90 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/s28_bench_$m.json 2> gpurun_out/s28_bench_$m.err
done
