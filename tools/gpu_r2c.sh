set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_gpu.py tests/test_api_errors_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err
tail -c 3000 gpurun_out/r2c_bench2.json; tail -5 gpurun_out/r2c_bench2.err
