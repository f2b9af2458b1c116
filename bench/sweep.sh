#!/bin/bash
# usage: bench/sweep.sh "<label>|<env assignments>|<bench args>" ...   one summary line per configuration
i=0
for spec in "$@"; do
  i=$((i+1))
  IFS='|' read -r label envs bargs <<< "$spec"
  env $envs python bench.py --steps 3 --no-cpu-baseline $bargs > gpurun_out/sw_$i.json 2>gpurun_out/sw_$i.err
  python - "$label" gpurun_out/sw_$i.json gpurun_out/sw_$i.err <<'PY'
import json, sys
label, jf, ef = sys.argv[1:4]
try:
    d = json.load(open(jf)); k = d["roofline"]["stage"]["kernels_ms"]
    print("%-28s value %6.1f G/s stage %6.2f ms e2e %6.2f ms | K1 %.2f K2 %.2f K3 %.2f K4 %.2f | frac K3 %.3f stage %.3f | ovf %.3f" % (
        label, d["value"] / 1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], k["partition_seq"], k["partition_fine"],
        k["probe_part"], k["unpermute"], d["roofline"]["frac"], d["roofline"]["stage"]["frac"], d["tables"]["overflow_frac"]))
except Exception as e:
    print(label, "FAILED", e); print(open(ef).read()[-600:])
PY
done
