set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3_tuning or partitioned_path_large" > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ai_pytest.log
timeout 600 python bench/k3_sweep.py --workload configs1 --steps 5 --out gpurun_out/r2ai_sweep_warp_rank.json k3_l2=7 k3_l2=1 k3_l2=2 k3_l2=4 k3_l2=9 k3_l2=10 k3_l2=5 k3_l2=0 > gpurun_out/r2ai_sweep.log 2>&1; tail -9 gpurun_out/r2ai_sweep.log
