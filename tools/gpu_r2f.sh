set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2f_pytest.log
W=5 K=20 bash tools/ab_local.sh 2>&1 | tee gpurun_out/r2f_ab_local.log
