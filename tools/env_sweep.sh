#!/bin/bash
# usage: tools/env_sweep.sh VAR v1 v2 ...   -- runs the cfg2 device bench once per value of the environment variable VAR
var=$1; shift
for v in "$@"; do
  env "$var=$v" python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$var=$v', round(d['value'],1), round(d['ms_per_step'],4), 'tc_ms', round(d['roofline']['tc_ms_per_step'],4), 'cc_ms', round(d['roofline']['cuda_core_conv_ms_per_step'],4))"
done
