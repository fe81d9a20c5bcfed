# final tree: full GPU suite + bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_final_tests.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_final_tests.txt
tail -4 gpurun_out/r02_final_tests.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_final_bench.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference arm rc=$?"; head -c 600 gpurun_out/r02_final_bench_reference.json; echo
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_final_bench.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('windows_ms'), 'lat', d['pipelining']['latency_ms_per_batch'])
P
