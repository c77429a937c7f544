#!/bin/bash
# Round-2 evidence run (one B200): c5job line, ncu launch list of the default bench command, ncu --set full of the
# fused SRC->EQ kernel (18944 clips x 1 s: all 148 CTAs busy), of the 4096-point FFT kernel (2368 clips x 10 s) and of
# the 2^16-point FFT kernel (C4 slice, 512 x 2^20 samples).
# Usage (through gpurun): bash tools/r2_profile.sh <tag>
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python bench.py --workload c5job --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-f64 > $out/${tag}_bench_c5job_n1.json 2> $out/${tag}_bench_c5job_n1.err
tail -c 1500 $out/${tag}_bench_c5job_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_ncu_launch_list_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f64 > $out/${tag}_ncu_launch_list.log 2>&1
echo "launch list rc=$?"
XZ_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:xz_mma_kernel --launch-skip 2 -c 1 -f \
    -o $out/${tag}_xz python tools/xz_perf.py 18944 1 44100 > $out/${tag}_ncu_xz.log 2>&1
echo "xz capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fft4096_r32_kernel --launch-skip 1 -c 1 -f \
    -o $out/${tag}_fft python bench.py --workload fft --clips 2368 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-f64 --no-parity > $out/${tag}_ncu_fft.log 2>&1
echo "fft capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fft65536_l32_kernel --launch-skip 2 -c 1 -f \
    -o $out/${tag}_fft_long python tools/bench_configs.py --only c4 --reps 2 > $out/${tag}_ncu_fft_long.log 2>&1
echo "long fft capture rc=$?"
ls -la $out
