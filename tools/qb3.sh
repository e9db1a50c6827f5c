#!/bin/bash
# usage (under gpurun): tools/qb3.sh <label> [bench.py args ...]  - bench line without the side runs, value / e2e only
label=$1; shift
timeout 600 python bench.py --no-cpu-baseline --no-other-configs "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', round(d['value']), round(d['e2e']['value']))"
