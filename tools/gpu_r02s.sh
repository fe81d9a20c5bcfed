# N1 producers: GPU parity, timing, ncu of the FP32 GEMM core; plus the --set full captures r02r missed (MANO, contact, level score)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_producers.py -m gpu -q -s > gpurun_out/r02s_tests.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02s_tests.txt
tail -12 gpurun_out/r02s_tests.txt
timeout 300 python tools/producers_bench.py 64 10 > gpurun_out/r02s_producers.json 2> gpurun_out/r02s_producers.err; echo "bench rc=$?"; cat gpurun_out/r02s_producers.json; tail -3 gpurun_out/r02s_producers.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_prod_launches.csv python tools/producers_bench.py 64 1 > gpurun_out/r02s_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/summarise_launches.py gpurun_out/r02s_prod_launches.csv > gpurun_out/r02s_prod_launches_summary.txt 2>&1; head -20 gpurun_out/r02s_prod_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_f32' --launch-skip 160 --launch-count 6 -o gpurun_out/r02s_gemm python tools/producers_bench.py 64 1 > gpurun_out/r02s_ncu_gemm.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mano_tc|k_obj_physics3|k_hand_level_score|k_hand_phys_score|k_feat_term' --launch-skip 40 --launch-count 16 -o gpurun_out/r02s_agg python tools/run_steps.py 3 > gpurun_out/r02s_ncu_agg.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summarize.py gpurun_out/r02s_ncu gpurun_out/r02s_gemm.ncu-rep gpurun_out/r02s_agg.ncu-rep > /dev/null
ncu -i gpurun_out/r02s_gemm.ncu-rep --page details > gpurun_out/r02s_gemm_details.txt 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -12
