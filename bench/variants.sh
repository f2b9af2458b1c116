#!/bin/bash
# sweep K3 variants on the default workload; prints one summary line per variant
for v in "$@"; do
  PK_K3_VARIANT=$v python bench.py --steps 3 --no-cpu-baseline > gpurun_out/b_v$v.json 2>gpurun_out/b_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_v$v.json"))
    k=d["roofline"]["stage"]["kernels_ms"]
    print("variant $v: value %.1f G/s  stage %.2f ms  e2e %.2f ms  K1 %.2f K2 %.2f K3 %.2f K4 %.2f  frac(K3) %.3f frac(stage) %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], k["partition_seq"], k["partition_fine"], k["probe_part"], k["unpermute"], d["roofline"]["frac"], d["roofline"]["stage"]["frac"]))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/b_v$v.err").read()[-800:])
PY
done
