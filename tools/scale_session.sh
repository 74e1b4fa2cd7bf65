#!/bin/bash
# Multi-GPU bench session (GPU box): tools/scale_session.sh <tag> "<gpus> <config> <scaling>" ...
# Each quoted triple runs bench.py once (torchrun for gpus > 1) and stores its JSON line in gpurun_out/<tag>_<config>_<scaling>_g<gpus>.json
tag=$1; shift
for spec in "$@"; do
    set -- $spec
    g=$1; c=$2; sc=$3
    out=gpurun_out/${tag}_${c}_${sc}_g${g}
    python bench.py --gpus $g --steps 4 --warmup 2 --config $c --scaling $sc --no-cpu-baseline > $out.json 2> $out.err
    echo "rc=$? $out: $(python -c "
import json,sys
try:
    d=json.loads([l for l in open('$out.json') if l.startswith('{')][-1])
    print('N=%d %.3f ms/step value %.4g e2e %.4g comm %.3f parity_ok=%s' % (d['config']['N'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('comm_ms_last_eval') or 0, (d.get('parity') or {}).get('ok')))
except Exception as e:
    print('no line', e)
")"
done
