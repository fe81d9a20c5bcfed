set -x
mkdir -p gpurun_out
for v in "" "--e2e-skip h2d" "--e2e-skip record" "--e2e-skip h2d,record" "--e2e-sets 5"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $v > gpurun_out/r02k.json 2> gpurun_out/r02k.err; echo "bench rc=$?"
python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/r02k.json')); print("VARIANT[%s]"%sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'lat', d['pipelining']['latency_ms_per_batch'])
PY
done
