mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mano_tc --launch-skip 2 --launch-count 1 -o gpurun_out/rep_mano python tools/mano_only.py 6400 > gpurun_out/r02q_ncu_mano.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/rep_mano.ncu-rep --page source --csv --print-source cuda > gpurun_out/r02q_mano_source.csv 2> gpurun_out/r02q_src.err; echo "src rc=$?"; wc -c gpurun_out/r02q_mano_source.csv
ncu -i gpurun_out/rep_mano.ncu-rep --page details > gpurun_out/r02q_mano_details.txt 2>&1
python tools/ncu_summarize.py gpurun_out/r02q_ncu gpurun_out/rep_mano.ncu-rep > /dev/null; rm -f gpurun_out/rep_mano.ncu-rep
ls -la gpurun_out
