#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain run first, then exactly ONE ncu pass of the same command
# (B200_PROFILING.md recipe; one ncu invocation per box call, always with a -k filter).
#   tools/profile_gpu.sh launches | k2 | k3 | k1
# Outputs land in gpurun_out/; tools/summarize_profiles.py digests them in the build container.
set -u
what=${1:-launches}
TAG=${2:-r02}
CMD="python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu-baseline --no-e2e --no-extras"
[ "$what" = "k1" ] && CMD="python bench.py --impl fft --steps 4 --warmup 3 --reps 1 --no-cpu-baseline --no-e2e --no-extras"
mkdir -p gpurun_out
timeout 200 $CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
case "$what" in
  launches)
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none \
        -k regex:"dft_i8_kernel|dft_seg_kernel|detect_kernel|hourly_kernel|stft_kernel" -c 40 --csv \
        --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 ;;
  k2) timeout 400 ncu --set full --clock-control none --import-source on -k regex:dft_i8_kernel -s 3 -c 1 -f \
        -o gpurun_out/${TAG}_k2_dft_i8 $CMD > gpurun_out/ncu_k2.log 2>&1 ;;
  k3) timeout 400 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 3 -c 1 -f \
        -o gpurun_out/${TAG}_k3_detect $CMD > gpurun_out/ncu_k3.log 2>&1 ;;
  k1) timeout 400 ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 3 -c 1 -f \
        -o gpurun_out/${TAG}_k1_stft_fft $CMD > gpurun_out/ncu_k1.log 2>&1 ;;
  seg) timeout 200 python tools/sweep_point.py 2048 0.5 seg > gpurun_out/prof_plain_seg.log 2>&1 || { echo "plain seg run failed"; exit 1; }
       timeout 400 ncu --set full --clock-control none --import-source on -k regex:dft_seg_kernel -s 2 -c 1 -f \
        -o gpurun_out/${TAG}_seg_2048_50 python tools/sweep_point.py 2048 0.5 seg 2 > gpurun_out/ncu_seg.log 2>&1 ;;
  *) echo "usage: $0 launches|k2|k3|k1|seg [tag]"; exit 2 ;;
esac
ls -la gpurun_out | tail -8
