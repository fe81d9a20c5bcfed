set -x
mkdir -p gpurun_out
timeout 600 python tools/pipeline_probe.py 20 > gpurun_out/r02h_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02h_probe.txt | tail -14
timeout 900 python -m pytest tests/test_e2e.py tests/test_sampler.py tests/test_evaluation.py -m gpu -q -x > gpurun_out/r02h_tests.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02h_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02h_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/r02h_bench_p0.json 2> gpurun_out/r02h_bench_p0.err; echo "bench p0 rc=$?"; tail -3 gpurun_out/r02h_bench_p0.err
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_k30.json 2> gpurun_out/r02h_bench_k30.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('r02h_bench','r02h_bench_p0','r02h_bench_k30'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['pipelining']['latency_ms_per_batch'], d['roofline']['avg_launch_ms'], d['per_step_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
