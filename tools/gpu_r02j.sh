set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bounds_build.py tests/test_force_optim.py tests/test_e2e.py -m gpu -q > gpurun_out/r02j_tests.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02j_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02j_bench.err
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_k30.json 2> gpurun_out/r02j_bench_k30.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('r02j_bench','r02j_bench_k30'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['pipelining']['latency_ms_per_batch'], d['roofline']['avg_launch_ms'], d['per_step_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
