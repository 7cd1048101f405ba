set -x
for v in idle64 idle200; do
for m in f16x3 f16; do
VLG_B200_LIB=scratch/variants_build/$v.so timeout 90 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/s19${v}_bench_$m.json 2> gpurun_out/s19${v}_bench_$m.err
done; done
