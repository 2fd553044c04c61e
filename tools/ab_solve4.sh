#!/bin/bash
# A/B of the solve variants on the same workload: mode 3 (default), mode 4 (bulk copy + PDL), mode 4 without PDL
cd $GRAFT_REPO_ROOT
for N in 55 30; do
for cfg in "ADMMB_SOLVE_MODE=4" "ADMMB_SOLVE_MODE=3"; do
  env $cfg python bench.py --cube $N --steps 20 --warmup 5 --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('N=$N %-40s value %7.1f  local %.3f rhs %.3f solve %.3f  step %.4f  solve GB/s %.0f' % ('$cfg', d['value'], p['local'], p['rhs'], p['solve'], p['step'], d['roofline_global']['achieved']))"
done; done
