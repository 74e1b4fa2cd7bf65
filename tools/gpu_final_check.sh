#!/bin/bash
# Round-end style check on the GPU box: parity suite, smoke, the default bench line (run from the repo root).
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.txt 2>&1; tail -3 gpurun_out/final_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps ${1:-10} --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; tail -c 600 gpurun_out/final_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench_1gpu.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "kernel_ms", "traffic")})
print("sph_sums", {k: d["sph_sums"][k] for k in ("ms", "frac_hbm", "frac_fp64")}, "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
print({k: v["ms"] for k, v in d["phases_last_eval"].items()})
PY
