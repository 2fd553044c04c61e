#!/bin/bash
# ncu passes for profiles/: launch list of ONE steady-state frame (10 ADMM iterations) + full captures of the two hot kernels.
# usage (on the GPU box): bash tools/profile_round.sh r2b [extra bench.py flags]
# bench.py --profile-region brackets its timed region with cudaProfilerStart/Stop and ncu runs with --profile-from-start off,
# so the 25 conditioning / warm-up frames run at full speed and only the profiled frame is serialised and replayed.
TAG=${1:-rX}; shift
export ADMMB_HOST_FACTOR=1   # keep cuSOLVER / cuBLAS setup kernels out of the launch numbering (the factor is the same)
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --cube 55 --steps 1 --warmup 5 --no-cpu-baseline --no-pairs --profile-region $*"
echo "bench command: $B" > $OUT/profile_${TAG}.log
ADMMB_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size --clock-control none \
    --csv --log-file $OUT/launches_${TAG}.csv $B >> $OUT/profile_${TAG}.log 2>&1
[ -n "$LIST_ONLY" ] && { tail -3 $OUT/profile_${TAG}.log; exit 0; }
ADMMB_NO_GRAPH=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_local_tets_hyper -s 4 -c 1 -f -o $OUT/prof_local_${TAG} \
    $B >> $OUT/profile_${TAG}.log 2>&1
[ -n "$LOCAL_ONLY" ] && { tail -3 $OUT/profile_${TAG}.log; exit 0; }
ADMMB_NO_GRAPH=1 ncu --profile-from-start off --set full --clock-control none -k regex:k_solve_level -s 104 -c 26 -f -o $OUT/prof_solve_${TAG} \
    $B >> $OUT/profile_${TAG}.log 2>&1
tail -3 $OUT/profile_${TAG}.log
