set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2d_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 6000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
