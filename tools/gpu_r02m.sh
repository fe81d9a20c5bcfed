set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_object_metrics.py tests/test_evaluation.py tests/test_bounds_build.py -m gpu -q > gpurun_out/r02m_tests.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02m_tests.txt
timeout 600 python tools/record_probe.py > gpurun_out/r02m_record.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02m_record.txt
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02m_bench.err
python - <<'PY'
import json
for f in ('r02m_bench',):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['pipelining']['latency_ms_per_batch'], d['roofline']['avg_launch_ms'], d['per_step_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
