#!/bin/bash
# quick A/B on the GPU box: scripts/quick.sh cfg3 tf32x3 f16x2 ...   (prints ms/step and the kernel shares)
wl=$1; shift
for eng in "$@"; do
  python bench.py --workload $wl --engine $eng --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_${wl}_${eng}.json 2> gpurun_out/q_${wl}_${eng}.err || tail -5 gpurun_out/q_${wl}_${eng}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/q_${wl}_${eng}.json"))
    r = d["roofline"]
    print("${wl}", d["engine"], "ms/step %.3f" % d["ms_per_step"], "e2e", round(d["e2e"]["value"]), r["kernel"], "%.3f ms" % r["avg_launch_ms"], "frac %.4f" % r["frac"], {k: round(v * d["ms_per_step"], 2) for k, v in r["share_of_step"].items()})
except Exception as ex:
    print("no result for ${wl} ${eng}:", ex)
PY
done
