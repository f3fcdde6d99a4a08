#!/bin/bash
# experiment helper: polygon device + e2e timings of the bench for each library variant in gpurun_variants/
cp myrenderer_b200/lib/libmyrenderer_b200.so /tmp/lib_default.so
for f in /tmp/lib_default.so gpurun_variants/*.so; do
  cp $f myrenderer_b200/lib/libmyrenderer_b200.so 2>/dev/null
  python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'star', round(d['polygons']['ms'],3), 'convex', round(d['polygons_convex']['ms'],3), 'large', round(d['polygons_large']['ms'],3), 'e2e poly ms', round(d['e2e']['ms_polygons'],3))"
done
cp /tmp/lib_default.so myrenderer_b200/lib/libmyrenderer_b200.so
