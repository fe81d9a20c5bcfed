set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sampler.py -m gpu -q -x > gpurun_out/r02c_tests_sampler.txt 2>&1; echo "sampler rc=$?"
tail -5 gpurun_out/r02c_tests_sampler.txt
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_sampler.py > gpurun_out/r02c_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_tests.txt
grep -E "passed|failed|FAILED|parity report" gpurun_out/r02c_tests.txt | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac']); print([ (s['stage'][:20], s['ms']) for s in d['stages']]); print(d['stages'][0]['kernels'])"
cap() { k=$1; skip=$2; cnt=$3; extra=$4
  timeout 300 ncu --set full --clock-control none $extra -k regex:$k --launch-skip $skip --launch-count $cnt -o gpurun_out/r02c_$k python tools/run_steps.py 2 > gpurun_out/r02c_ncu_$k.log 2>&1; echo "ncu $k rc=$?"; }
cap k_head_tc 30 1 "--import-source on"
cap k_pose_tc 30 1 "--import-source on"
cap mano_forward_kernel 5 4 ""
cap k_obj_physics3 1 1 ""
cap k_hand_level_score 4 4 ""
cap k_hand_level_fuse 4 1 ""
cap k_hand_phys_score 1 1 ""
cap k_obj_final 1 1 ""
cap k_force_anchors 2 2 ""
cap k_post_step 3 1 ""
cap k_postprocess_hand 2 2 ""
ls -la gpurun_out; du -sh gpurun_out
