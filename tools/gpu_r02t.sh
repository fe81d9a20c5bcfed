# N1 producers on tensor cores: parity of both paths, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_producers.py -m gpu -q -s -x > gpurun_out/r02t_tests.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02t_tests.txt
tail -30 gpurun_out/r02t_tests.txt
timeout 300 python tools/producers_bench.py 64 10 > gpurun_out/r02t_producers_tc.json 2> gpurun_out/r02t_producers_tc.err; echo "bench rc=$?"; cat gpurun_out/r02t_producers_tc.json; tail -3 gpurun_out/r02t_producers_tc.err
timeout 300 python tools/producers_bench.py 64 5 strict > gpurun_out/r02t_producers_simt.json 2>&1; cat gpurun_out/r02t_producers_simt.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02t_prod_launches.csv python tools/producers_bench.py 64 1 > gpurun_out/r02t_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/summarise_launches.py gpurun_out/r02t_prod_launches.csv > gpurun_out/r02t_prod_launches_summary.txt 2>&1; head -24 gpurun_out/r02t_prod_launches_summary.txt
