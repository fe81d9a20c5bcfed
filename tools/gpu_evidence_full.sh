# Round-2 evidence run: GPU tests, bench (both arms), ncu launch list, ncu --set full of the kernels >= 3 % of the step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_tests.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_tests.txt
tail -4 gpurun_out/r02_tests.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench.err
head -c 1200 gpurun_out/r02_bench.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/r02_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/summarise_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1; head -30 gpurun_out/r02_launches_summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mano_tc|mano_forward_kernel|k_pose_tc|k_obj_physics3|k_hand_level_score|k_head_tc|k_reduce|k_post_step|k_hand_phys_score|k_postprocess_hand' --launch-skip 200 --launch-count 50 -o gpurun_out/r02_full python tools/run_steps.py 3 > gpurun_out/r02_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summarize.py gpurun_out/r02_ncu gpurun_out/r02_full.ncu-rep > /dev/null; rm -f gpurun_out/r02_full.ncu-rep
ls -la gpurun_out
timeout 120 python tools/diag_pose_timeline.py > gpurun_out/r02_pose_timeline.txt 2>&1; echo "timeline rc=$?"
