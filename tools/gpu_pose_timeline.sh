# pose-encoder timeline at the headline shape (stamped twin of the library) + sampler parity + bench; tight timeouts
mkdir -p gpurun_out
timeout 90 python tools/debug_pose_eval.py 2>&1 | grep -v bad; echo "eval rc=$?"
timeout 90 python tools/diag_pose_timeline.py > gpurun_out/r02v_pose_timeline.txt 2>&1; echo "rc=$?"; cat gpurun_out/r02v_pose_timeline.txt | cut -c1-700
timeout 120 python -m pytest tests/test_sampler.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02v_bench.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'lat', d['pipelining']['latency_ms_per_batch'])
print(json.dumps(d['stages'][0]['kernels']), d['roofline']['frac'], d['clocks'])
P
