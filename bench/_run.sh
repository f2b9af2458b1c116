set -x
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -s > gpurun_out/r2w_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2w_configs1_n2.json 2> gpurun_out/r2w_configs1_n2.err; echo "configs1 n2 rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r2w_configs1_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N2 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['e2e']['last_call_ms'],'files',d['e2e_files']['ms_per_step'],d['e2e_files']['last_call_ms'],'exch',d['exchange'],[ (r['kernels_ms']['k_probe_ms'], r['group_table_bytes']) for r in d['per_rank']])
P
