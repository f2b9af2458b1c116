set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reduce_device_wide or k3_tuning or full_size or partitioned_path_large" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_pytest.log
for nb in 0 3 4; do
timeout 600 python bench.py --no-cpu-baseline --no-ncu --e2e-batches $nb > gpurun_out/r2z_bench_b$nb.json 2> gpurun_out/r2z_bench_b$nb.err; python - $nb <<'P'
import json,sys
for l in open('gpurun_out/r2z_bench_b%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('batches',sys.argv[1],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['e2e']['device_timeline_ms'],'files',d['e2e_files']['ms_per_step'])
P
done
