# N1 producers: final parity of both CUDA paths + the chained predict_from_features test, bench --config 6, ncu of the tcgen05 GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_producers.py -m gpu -q -s > gpurun_out/r02u_tests.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02u_tests.txt
tail -25 gpurun_out/r02u_tests.txt | cut -c1-400
timeout 600 python bench.py --config 6 --steps 10 --warmup 3 > gpurun_out/r02u_bench_config6.json 2> gpurun_out/r02u_bench_config6.err; echo "bench rc=$?"; cat gpurun_out/r02u_bench_config6.json | cut -c1-1800; tail -3 gpurun_out/r02u_bench_config6.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_prod_launches.csv python tools/producers_bench.py 64 1 > gpurun_out/r02u_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/summarise_launches.py gpurun_out/r02u_prod_launches.csv > gpurun_out/r02u_prod_launches_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc' --launch-skip 262 --launch-count 40 -o gpurun_out/r02u_gemm python tools/producers_bench.py 64 1 > gpurun_out/r02u_ncu_gemm.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summarize.py gpurun_out/r02u_ncu_producers gpurun_out/r02u_gemm.ncu-rep > /dev/null
rm -f gpurun_out/*.ncu-rep
