#!/bin/bash
# experiment helper: run the polygon part of the bench with each library variant in gpurun_variants/
cp myrenderer_b200/lib/libmyrenderer_b200.so /tmp/lib_default.so
for f in /tmp/lib_default.so gpurun_variants/*.so; do
  cp $f myrenderer_b200/lib/libmyrenderer_b200.so 2>/dev/null
  python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['polygons']['ms'],3), round(d['polygons']['value']/1e6,2), round(d['polygons_convex']['ms'],3), round(d['polygons_convex']['value']/1e6,2), d.get('polygon_tiers'))"
done
cp /tmp/lib_default.so myrenderer_b200/lib/libmyrenderer_b200.so
