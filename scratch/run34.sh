set -x
for v in tb1 tb2 tb4; do
VLG_B200_LIB=scratch/variants_build/$v.so timeout 200 python bench.py --config 5 --precision f16 --no-cpu --no-other > gpurun_out/s24${v}_c5_f16.json 2> gpurun_out/s24${v}_c5_f16.err
done
