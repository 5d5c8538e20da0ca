#!/bin/bash
# build library variants in parallel, then run the c3 bench with each: tools/variant_bench.sh "name:flags" ...
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; fl=${v#*:}
  ( SCB_LIB_SUFFIX=_$name SCB_EXTRA_FLAGS="$fl" python sparsify_clip_b200/build.py --force > gpurun_out/build_$name.log 2>&1 ) &
done
wait
for v in "$@"; do
  name=${v%%:*}
  SCB_LIB_SUFFIX=_$name timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('[$name]', round(d['ms_per_step'],3), 'ms frac', round(r['frac'],4), {k:round(v,3) for k,v in r['sweeps_only']['per_pass_ms'].items()}, d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['ms_per_step'],3))"
done
