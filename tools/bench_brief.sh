#!/bin/bash
# quick GPU check: parity tests + short bench with the phase split (run on the GPU box)
T=$(python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
python bench.py --steps ${1:-5} --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_brief.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_brief.json").read())
print("value %.4g p-s/s  %.2f ms/step  e2e %.4g  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
print({k: v["ms"] for k, v in d["phases_last_eval"].items()}, "knn_retries", d.get("knn_retries"))
PY
echo "pytest -m gpu: $T"
