#!/bin/bash
# experiment: polygon kernel knobs (MR_TUNE bit 0 = no items, bits 1-4 = item refresh period, bits 8-12 = serial-select threshold)
for t in 0x000 0x001 0x800 0x801 0x804 0x806 0x808 0x1008 0x1001 0x400; do
  MR_TUNE=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$t', round(d['polygons']['ms'],3), round(d['polygons']['value']/1e6,2))"
done
