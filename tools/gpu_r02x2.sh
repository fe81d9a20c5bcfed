mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_e2e.py tests/test_headline_parity.py tests/test_evaluation.py -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02x_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02x_bench.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], d['value_windows_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('windows_ms'), 'lat', d['pipelining']['latency_ms_per_batch'], 'launches', d['gpu_launches'])
print(d['per_step_ms'])
P
