set -x
python scratch/prof_tc.py 148 2 f16 > gpurun_out/r2h_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_curve_kernel -s 2 -c 1 -o gpurun_out/r2h_tc_f16 python scratch/prof_tc.py 148 2 f16 > gpurun_out/r2h_ncu.log 2>&1
