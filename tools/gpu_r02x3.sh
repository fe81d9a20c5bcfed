timeout 150 python tools/stall_probe.py 400 2>&1 | grep "^i=\|^batches\|slow iter" | cut -c1-400
for cfg in "2 1 4" "1 1 4"; do
set -- $cfg
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --agg-slots $1 --side-slots $2 --e2e-sets $3 > /tmp/b.json 2>/dev/null
python - <<P
import json
d=json.load(open('/tmp/b.json'))
print('agg $1 side $2 sets $3: value', d['ms_per_step'], d['value_windows_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['windows_ms'], 'last', d['per_step_ms'][-1], 'lat', d['pipelining']['latency_ms_per_batch'])
print(d['per_step_ms'])
P
done
