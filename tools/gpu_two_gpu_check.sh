# 2-GPU sanity of the final tree: smoke(), torchrun bench at N=2, reference arm
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_n2_bench_n2.json 2> gpurun_out/r02_n2_bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/r02_n2_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_n2_bench_n2.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('windows_ms'))
P
