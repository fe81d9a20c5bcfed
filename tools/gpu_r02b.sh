set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_parity.py tests/test_object_metrics.py tests/test_evaluation.py -m gpu -q -s -k "random or metrics or evaluation" > gpurun_out/r02b_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests.txt
tail -30 gpurun_out/r02b_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02b_bench.err
timeout 600 python bench.py --config 5 --steps 5 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; echo "bench5 rc=$?"
tail -3 gpurun_out/r02b_bench_c5.err
for k in mano_forward_kernel k_pose_tc k_stage_x k_obj_physics3 k_hand_level_score k_head_tc k_hand_phys_score; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 12 --launch-count 1 -o gpurun_out/r02b_$k python tools/run_steps.py 2 > gpurun_out/r02b_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out
