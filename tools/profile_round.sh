#!/bin/bash
# ncu passes for profiles/: launch list of 3 steady-state ADMM iterations + full captures of the two hot kernels.
# usage (on the GPU box): bash tools/profile_round.sh r1d
TAG=${1:-rX}
export ADMMB_HOST_FACTOR=1   # keep cuSOLVER / cuBLAS setup kernels out of the launch numbering (the factor is the same)
OUT=gpurun_out
mkdir -p $OUT
NL=$(python - <<'PY'
import sys, os
sys.path.insert(0, "admm-elastic-sca_b200/pyhost")
import admm_b200, scenes
sc = scenes.cube_scene(55)
s = admm_b200.System(sc)
print(s.info()["n_levels"])
PY
)
PER=$((2 + 2 * NL))                    # local + rhs + 2 x levels (ADMMB_NO_GRAPH=1: one launch per kernel)
SKIP=$((23 * (10 * PER + 2) + 4))      # 20 conditioning + 3 warm-up frames (+ frame begin / end), upload permutes
echo "levels $NL, launches per iteration $PER, skipping $SKIP" > $OUT/profile_${TAG}.log
ADMMB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size --clock-control none \
    -s $SKIP -c $((3 * PER + 2)) --csv --log-file $OUT/launches_${TAG}.csv python bench.py --cube 55 --steps 2 --warmup 3 --no-cpu-baseline --no-pairs >> $OUT/profile_${TAG}.log 2>&1
[ -n "$LIST_ONLY" ] && { tail -3 $OUT/profile_${TAG}.log; exit 0; }
ADMMB_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:k_local_tets_hyper -s 235 -c 1 -f -o $OUT/prof_local_${TAG} \
    python bench.py --cube 55 --steps 2 --warmup 3 --no-cpu-baseline --no-pairs >> $OUT/profile_${TAG}.log 2>&1
ADMMB_NO_GRAPH=1 ncu --set full --clock-control none -k regex:k_solve_level -s $((23 * 10 * 2 * NL)) -c $((2 * NL)) -f -o $OUT/prof_solve_${TAG} \
    python bench.py --cube 55 --steps 2 --warmup 3 --no-cpu-baseline --no-pairs >> $OUT/profile_${TAG}.log 2>&1
tail -3 $OUT/profile_${TAG}.log
