set -x
bash scratch/run_quick.sh s3
for m in f16x3 f16; do
VLG_B200_LIB=scratch/variants_build/noahead.so timeout 300 python bench.py --steps 10 --warmup 3 --precision $m --no-cpu --no-other > gpurun_out/s3na_bench_$m.json 2> gpurun_out/s3na_bench_$m.err
done
