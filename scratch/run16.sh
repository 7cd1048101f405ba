set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not config3_final and not full_config1 and not full_pair_list" 2>&1 | tail -5 > gpurun_out/s4_pytest.log
for v in default noahead fwd bwd; do
for m in f16x3 f16; do
lib=scratch/variants_build/$v.so; [ $v = default ] && lib=vae-latent-geometry_b200/libvlg_b200.so
VLG_B200_LIB=$lib timeout 90 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/s4${v}_bench_$m.json 2> gpurun_out/s4${v}_bench_$m.err
done; done
