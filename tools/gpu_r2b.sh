set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.log
W=3 K=5 bash tools/ab_local.sh 2>&1 | tee gpurun_out/r2b_ab_local.log
