#!/bin/bash
# usage (under gpurun): tools/qb2.sh <label> [bench.py args ...]  - full-length bench line incl. other_configs, value / e2e only
label=$1; shift
timeout 600 python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', round(d['value']), round(d['e2e']['value']), {k[:40]:(round(v['value']),round(v['e2e']['value'])) for k,v in d.get('other_configs',{}).items()})"
