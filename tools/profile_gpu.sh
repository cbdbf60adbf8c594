#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain run first, then the ncu launch list and one --set full
# capture per kernel of the same command (B200_PROFILING.md recipe).  Outputs land in gpurun_out/.
set -u
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dft_i8_kernel|detect_kernel|hourly_kernel|stft_kernel" \
    -c 40 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dft_i8_kernel -s 3 -c 1 -o gpurun_out/r01_k2_dft_i8 $CMD > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 3 -c 1 -o gpurun_out/r01_k3_detect $CMD > gpurun_out/ncu_k3.log 2>&1
CMD1="python bench.py --impl fft --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD1 > gpurun_out/prof_plain_fft.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 3 -c 1 -o gpurun_out/r01_k1_stft_fft $CMD1 > gpurun_out/ncu_k1.log 2>&1
ls -la gpurun_out
