#!/bin/bash
# usage (under gpurun): tools/variant_bench.sh <name> [<name> ...]  - quick bench of devtools_build/liblyra_b200_<name>.so variants ("product" = the built library)
for v in "$@"; do
  lib=devtools_build/liblyra_b200_$v.so
  [ "$v" = product ] && lib=lyra_b200/liblyra_b200.so
  LYRA_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --no-other-configs --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', round(d['value']), round(d['e2e']['value']), {k[:8]+k[-1]:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})"
done
