#!/bin/bash
# mesh-size sweep of DESIGN.md section 7 (one B200, NeoHookean cube, direct solver)
cd $GRAFT_REPO_ROOT
for N in 30 55 70 87 110; do
  python bench.py --cube $N --steps 10 --warmup 5 --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; s=d['setup']; r=d['roofline']; g=d['roofline_global']
print('N=$N value %.1f e2e %.1f | step %.3f ms = local %.3f rhs %.3f solve %.3f | solve %.0f GB/s (%.2f) | local alg %.2f TF/s frac %.3f unfused %.3f flops/tet %.0f | nnzL %.1fM factor %.2f GB levels %d setup %.1f s' % (
 d['value'], d['e2e']['value'], p['step'], p['local'], p['rhs'], p['solve'], g['achieved'], g['frac'], r['achieved'], r['frac'], r['frac_of_unfused_peak'], r['algorithmic_flops_per_tet_iteration'], s['nnz_L']/1e6, s['factor_bytes']/1e9, s['levels'], s['seconds']))"
done
