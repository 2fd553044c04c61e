cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2u_pytest.log
ADMMB_NO_SPLIT=1 python bench.py --cube 55 --steps 20 --warmup 5 --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('%-40s value %7.1f  local %.3f rhs %.3f solve %.3f' % ('no-split', d['value'], p['local'], p['rhs'], p['solve']))"
W=5 K=20 bash tools/ab_local.sh 2>&1 | tee gpurun_out/r2u_ab_local.log
