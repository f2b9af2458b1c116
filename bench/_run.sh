set -x
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu > gpurun_out/r2al_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2al_pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2al_configs1_n2.json 2> gpurun_out/r2al_configs1_n2.err; echo "configs1 n2 rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r2al_configs1_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N2 ms',d['ms_per_step'],'G/s',d['value']/1e9,'exch',{k:v for k,v in d['exchange'].items() if k!='what'}, d['step_ms'])
P
tail -3 gpurun_out/r2al_configs1_n2.err
