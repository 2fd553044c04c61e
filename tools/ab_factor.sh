#!/bin/bash
# setup time: host-only factorisation vs large fronts on the device, per-level log on stderr
IFS=","; for cfg in ${CFGS:-55 1 1536,55 0 1536,55 0 2560,87 0 1536,87 0 3000}; do
  IFS=" "; set -- $cfg
  ADMMB_FACTOR_VERBOSE=1 ADMMB_HOST_FACTOR=$2 ADMMB_GPU_FRONT_MIN=$3 python bench.py --cube $1 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/factor_$1_$2_$3.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['setup']; print('cube=$1 host_only=$2 min_front=$3 setup %.2fs factor %.2fs value %.1f' % (s['seconds'], s['factor_seconds'], d['value']))"
  grep "setup\]\|one-at" gpurun_out/factor_$1_$2_$3.err
done
