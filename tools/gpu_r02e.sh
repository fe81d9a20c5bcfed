set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mano.py tests/test_sampler.py -m gpu -q -x > gpurun_out/r02e_tests_first.txt 2>&1; echo "first rc=$?"
tail -15 gpurun_out/r02e_tests_first.txt
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_sampler.py --deselect tests/test_mano.py > gpurun_out/r02e_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_tests.txt
grep -E "passed|failed|FAILED|parity report \(" gpurun_out/r02e_tests.txt | cut -c1-400 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac']); print([ (s['stage'][:20], s['ms'], s.get('frac')) for s in d['stages']]); print(d['stages'][0]['kernels'])"
timeout 600 python bench.py --config 3 --steps 5 > gpurun_out/r02e_bench_c3.json 2> gpurun_out/r02e_bench_c3.err; echo "bench3 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_c3.json')); print(d['roofline']); print([r for r in d['results'] if r['case']=='a_lbs_verts'])"
cap() { k=$1; skip=$2; cnt=$3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count $cnt -o gpurun_out/rep_$k python tools/run_steps.py 2 > gpurun_out/r02e_ncu_$k.log 2>&1; echo "ncu $k rc=$?"; }
cap k_head_tc 45 2
cap k_pose_tc 45 2
cap k_mano_tc 5 4
cap k_obj_physics3 1 1
cap k_hand_level_score 4 4
cap k_hand_level_fuse 4 1
cap k_hand_phys_score 1 1
cap k_obj_final 1 1
cap k_force_anchors 2 2
cap k_post_step 2 3
cap k_postprocess_hand 2 2
python tools/ncu_summarize.py gpurun_out/r02e_ncu gpurun_out/rep_*.ncu-rep > /dev/null
mkdir -p /tmp/keep && mv gpurun_out/rep_k_head_tc.ncu-rep gpurun_out/rep_k_pose_tc.ncu-rep /tmp/keep/ ; rm -f gpurun_out/rep_*.ncu-rep; mv /tmp/keep/*.ncu-rep gpurun_out/
ls -la gpurun_out; du -sh gpurun_out
