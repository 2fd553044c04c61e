cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_host_cpp.py tests/test_dropin_scene_layer_gpu.py tests/test_forcebuilder_batched.py -x -q -s 2>&1 | grep -E "bit-identical|passed|failed|Error|error|SoA" | tail -40 | tee gpurun_out/r2w_pytest.log
python tools/trial_stats.py 55 2>&1 | tail -8 | tee gpurun_out/r2w_trials55.log
