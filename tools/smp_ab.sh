for s in 1 2 3 4 6; do echo "== STC_DBG=$s"; B200_STC_DBG=$s timeout 300 python bench.py --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], {k:round(v,4) for k,v in d['stage_ms'].items() if 'gated' not in k and 'round1' not in k}, d['map'])
"; done
