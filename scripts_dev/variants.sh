#!/bin/bash
# A/B of library builds on one box: isolated pass time (dsym_bench.py) and in-loop pass time (bench.py) per variant
for lib in "" $(ls scripts_dev/lib_*.so 2>/dev/null); do
  if [ -z "$lib" ]; then export -n SGV_LIB; unset SGV_LIB; name=default; else export SGV_LIB=$PWD/$lib; name=$(basename $lib .so); fi
  echo "=== $name"
  timeout 200 python scripts_dev/dsym_bench.py 2>&1 | tail -1
  timeout 300 python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('in-loop avg_launch_ms %.4f frac %.4f iso_ms %.4f value %.1f it/s launches %d' % (r['avg_launch_ms'], r['frac'], r['isolated_launch_ms'], d['value'], d['gpu_launches']))"
done
