set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bgzf or anchor_dir or k3_tuning or pinned or cli or edge_cases or sharded_anchorer_world1" > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ac_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ac_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-ncu > gpurun_out/r2ac_bench_configs1.json 2> gpurun_out/r2ac_bench_configs1.err
for i in 1 2; do timeout 600 python bench.py --index-e2e configs1 --no-cpu-baseline > gpurun_out/r2ac_index_configs1_$i.json 2> gpurun_out/r2ac_index_$i.err; done
python - <<'P'
import json
for l in open('gpurun_out/r2ac_bench_configs1.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'files',d['e2e_files']['ms_per_step'])
for i in (1,2):
    for l in open('gpurun_out/r2ac_index_configs1_%d.json'%i):
        if l.startswith('{'):
            d=json.loads(l); print('index',d['total_s'],d['log'])
P
