set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "
import sys; sys.path.insert(0,'admm-elastic-sca_b200/pyhost')
import admm_b200, json
print('FP64PROBE', json.dumps(admm_b200.probe_fp64()))
m,f = admm_b200.fastmath_selftest(samples=1<<28, seed=7)
print('FASTMATH mism', m.tolist(), 'fallbacks', f.tolist())
" 2>&1 | tee gpurun_out/r2a_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log
W=20 K=5 bash tools/ab_local.sh 2>&1 | tee gpurun_out/r2a_ab_local.log
