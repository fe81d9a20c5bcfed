mkdir -p gpurun_out
for i in 1 2 3 4; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_$i.json 2> gpurun_out/r02p_bench_$i.err; echo "bench rc=$?"; tail -3 gpurun_out/r02p_bench_$i.err
python - $i <<'PY'
import json,sys
d=json.load(open('gpurun_out/r02p_bench_%s.json'%sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']); print(d['per_step_ms']); print(d['host_enqueue_wait_ms'])
PY
done
