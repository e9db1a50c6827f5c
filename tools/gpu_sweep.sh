#!/bin/bash
# dev helper run under gpurun: quick parity + timing for both tile sizes
mkdir -p gpurun_out
for S in 8 16; do
  echo "=== TILE_STREAMS=$S"
  LYRA_B200_TILE_STREAMS=$S python tools/gpu_probe.py 4096 3 2>&1 | grep -E "parity|device-resident|MISMATCH"
  LYRA_B200_TILE_STREAMS=$S python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f frames/s  ms/step %.3f  e2e %.0f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))
for k,v in d['roofline']['kernels'].items(): print('   %-18s %.4f ms' % (k, v['ms_per_launch']))
"
done
