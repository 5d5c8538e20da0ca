#!/bin/bash
# build tuning variants of the library in parallel and time the sweeps of each: tools/variants.sh "name:flags" ...
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; fl=${v#*:}
  ( SCB_LIB_SUFFIX=_$name SCB_EXTRA_FLAGS="$fl" python sparsify_clip_b200/build.py --force > gpurun_out/build_$name.log 2>&1 ) &
done
wait
for v in "$@"; do
  name=${v%%:*}
  SCB_LIB_SUFFIX=_$name timeout 120 python tools/sweep_time.py 7 2>&1 | grep -v Warning
done
