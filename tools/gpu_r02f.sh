set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_tests.txt
grep -E "passed|failed|FAILED|parity report \(|pytest rc" gpurun_out/r02f_tests.txt | cut -c1-400 | tail -14
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02f_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --record-stream side --no-cpu-baseline > gpurun_out/r02f_bench_side.json 2> gpurun_out/r02f_bench_side.err; echo "bench side rc=$?"
python - <<'PY'
import json
for f in ('r02f_bench','r02f_bench_side'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
for c in 1 3 5; do timeout 600 python bench.py --config $c --steps 5 > gpurun_out/r02f_bench_c$c.json 2> gpurun_out/r02f_bench_c$c.err; echo "bench c$c rc=$?"; cut -c1-600 gpurun_out/r02f_bench_c$c.json; done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "ref rc=$?"; cut -c1-800 gpurun_out/r02f_bench_ref.json
ls -la gpurun_out; du -sh gpurun_out
