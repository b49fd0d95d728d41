# same-box A/B of the certified stretches in the host loop (DQ_NO_CERTS=1 = walk and subtract everything)
for p in "" 1 "" 1; do
  echo "== DQ_NO_CERTS=$p"
  env ${p:+DQ_NO_CERTS=1} timeout 100 python scripts/trace_e2e.py 2>&1 | grep -E "^call|host loop done" | tail -4
done
for shape in "" "0,1"; do
for p in "" 1; do
  echo "== bench DQ_NO_CERTS=$p DQ_HOST_THREADS=$shape"
  env ${p:+DQ_NO_CERTS=1} ${shape:+DQ_HOST_THREADS=$shape} timeout 120 python bench.py --no-extras --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'pageable',round(d['e2e']['pageable_buffers']['value']))
"
done
done
