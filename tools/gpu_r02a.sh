set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=15 -s > gpurun_out/r02a_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_tests.txt
echo skip bench
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mano_forward_kernel|k_pose_tc|k_stage_x|k_obj_physics3|k_hand_level_score|k_head_tc|k_reduce|k_post_step|k_hand_phys_score|k_postprocess_hand' --launch-skip 260 --launch-count 60 -o gpurun_out/r02a_full python tools/run_steps.py 3 > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r02a_tests.txt
cat gpurun_out/r02a_bench.json | head -c 1500
