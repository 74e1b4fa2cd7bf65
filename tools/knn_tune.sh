#!/bin/bash
# kNN variants side by side (GPU box): bash tools/knn_tune.sh "ENV=V,ENV=V" ...   ("-" = defaults)
for cfg in "$@"; do
  envs=""; [ "$cfg" != "-" ] && envs=$(echo "$cfg" | tr ',' ' ')
  env $envs python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/knn_tune.json
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("/tmp/knn_tune.json"))
p = d["phases_last_eval"]
print("%-40s step %.2f ms  knn %.3f  grav %.3f  total/eval %.3f  retries %s" % (sys.argv[1], d["ms_per_step"], p["knn"]["ms"], p["gravity"]["ms"], p["total"]["ms"], d.get("knn_retries")), flush=True)
PY
done
