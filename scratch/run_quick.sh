# quick check of a kernel change: short parity tests, the two headline benches, phase statistics
set -x
tag=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not config3_final and not full_config1 and not full_pair_list" 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision f16x3 --no-cpu --no-other > gpurun_out/${tag}_bench_f16x3.json 2> gpurun_out/${tag}_bench_f16x3.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision f16 --no-cpu --no-other > gpurun_out/${tag}_bench_f16.json 2> gpurun_out/${tag}_bench_f16.err
if [ -f scratch/variants_build/stats.so ]; then
timeout 300 python scratch/stats_tc.py scratch/variants_build/stats.so f16x3 > gpurun_out/${tag}_stats_f16x3.log 2>&1
timeout 300 python scratch/stats_tc.py scratch/variants_build/stats.so f16 > gpurun_out/${tag}_stats_f16.log 2>&1
fi
