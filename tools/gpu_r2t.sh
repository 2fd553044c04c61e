cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py tests/test_api_errors_gpu.py -m gpu -x -q 2>&1 | tail -4
for N in 30 55 110; do
python bench.py --cube $N --solver pcg --steps 5 --warmup 3 --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']
print('PCG N=$N value %.1f | local %.3f rhs %.3f solve %.3f | cg its/admm it %s' % (d['value'], p['local'], p['rhs'], p['solve'], d['config'].get('cg_iterations_per_admm_iteration')))"
done
