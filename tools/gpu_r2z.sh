cd $GRAFT_REPO_ROOT
timeout 400 python tools/strong_scaling.py --cube 110 --solver pcg 2>&1 | grep value | tee gpurun_out/r2z_pcg110_1gpu.log
timeout 200 python tools/strong_scaling.py --cube 55 --solver pcg --steps 20 2>&1 | grep value | tee -a gpurun_out/r2z_pcg110_1gpu.log
timeout 900 bash tools/profile_round.sh r2z 2>&1 | tail -3
