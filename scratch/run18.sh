set -x
for v in nofh nofwd nofh_single; do
for m in f16x3 f16; do
VLG_B200_LIB=scratch/variants_build/$v.so timeout 90 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/s6${v}_bench_$m.json 2> gpurun_out/s6${v}_bench_$m.err
done; done
