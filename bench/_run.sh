set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3_tuning" > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2af_pytest.log
timeout 600 python bench/k3_sweep.py --workload configs1 --steps 5 --out gpurun_out/r2af_sweep_k3_l2_direct.json k3_l2=7 k3_l2=9 k3_l2=10 k3_l2=11 k3_l2=12 k3_l2=13 k3_l2=14 k3_l2=7 > gpurun_out/r2af_sweep.log 2>&1; tail -9 gpurun_out/r2af_sweep.log
