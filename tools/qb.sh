#!/bin/bash
# usage (under gpurun): tools/qb.sh <label> [bench.py args ...]   - one quick bench line; LYRA_B200_LIB / priorities etc. come from the environment
label=$1; shift
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --steps 8 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', round(d['value']), round(d['e2e']['value']), d['config']['output_checksum'], {k[:8]+k[-1]:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})"
