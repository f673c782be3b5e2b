#!/bin/bash
# GPU-box helper: GPU tests, then compact bench summaries for several workloads.
#   tools/gpu_check.sh [--notest] "C2:50 C3:200 C5:250"   (workload:loci-per-gpu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" == "--notest" ]; then shift; else timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
for spec in ${1:-C2:50}; do
  w=${spec%%:*}; n=${spec##*:}
  timeout 600 python bench.py --workload $w --loci $n --steps ${STEPS:-100} --warmup 5 --no-cpu --no-extra > gpurun_out/bench_${w}_${n}.json 2> gpurun_out/bench_${w}_${n}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${w}_${n}.json")); r=d["roofline"]
    print("${w}:${n}", "ms/step %.4f"%d["ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], "kernel_us %.2f"%r["us_per_launch"],
          "GB/s %.0f"%r["achieved"], "frac %.3f"%r["frac"], "survey_frac %.3f"%r["survey_formula_frac"],
          "kernel_units/s %.3g"%r["kernel_partials_per_sec"], "value %.3g"%d["value"], "clk", d["clocks"]["sm_mhz"])
except Exception as ex:
    print("${w}:${n} FAILED", ex); print(open("gpurun_out/bench_${w}_${n}.err").read()[-1500:])
PY
done
