run() { python bench.py --no-decode --no-cpu-baseline --steps 20 > /tmp/b.json 2>/tmp/b.err; python -c "
import json
d=json.loads([l for l in open('/tmp/b.json') if l.startswith('{')][-1])
print('$1', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['breakdown'].items()})
"; }
