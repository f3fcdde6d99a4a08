#!/bin/bash
# experiment helper: time scripts/tune_large.py with each library variant in gpurun_variants/
cp myrenderer_b200/lib/libmyrenderer_b200.so /tmp/lib_default.so
for f in /tmp/lib_default.so gpurun_variants/*.so; do
  cp $f myrenderer_b200/lib/libmyrenderer_b200.so 2>/dev/null
  echo "$f $(python scripts/tune_large.py ${1:-log} ${2:-100000} 2>&1 | tail -1)"
done
cp /tmp/lib_default.so myrenderer_b200/lib/libmyrenderer_b200.so
