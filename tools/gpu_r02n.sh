set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02n_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_tests.txt
grep -E "passed|failed|FAILED|parity report \(|pytest rc" gpurun_out/r02n_tests.txt | cut -c1-300 | tail -14
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02n_bench.err
timeout 600 python bench.py --config 3 --steps 5 > gpurun_out/r02n_bench_c3.json 2> gpurun_out/r02n_bench_c3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['pipelining']['latency_ms_per_batch'], d['roofline']['avg_launch_ms'], d['per_step_ms'])
print([(s['stage'][:24], s['ms'], s.get('frac')) for s in d['stages']])
d=json.load(open('gpurun_out/r02n_bench_c3.json')); print(d['roofline']); print([r for r in d['results'] if r['case']=='a_lbs_verts'])
PY
