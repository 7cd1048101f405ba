set -x
tag=${1:-q}
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not config3_final and not full_config1 and not full_pair_list" 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
for m in f16x3f f16; do
timeout 90 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/${tag}_bench_$m.json 2> gpurun_out/${tag}_bench_$m.err
done
