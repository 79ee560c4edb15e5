for s in 1 0; do echo "== SEL_STAGE=$s"; B200_SEL_STAGE=$s timeout 300 python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], {k:round(v,4) for k,v in d['stage_ms'].items() if k in ('sample_hist','bound','select','rank')}, d['map'])
        c=d.get('scaleout')
        if c: print('c5', c['ms_per_step'], {k:round(v,4) for k,v in c['stage_ms'].items() if k in ('sample_hist','bound','select','rank')}, c['map'])
"; done
