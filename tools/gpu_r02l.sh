mkdir -p gpurun_out
timeout 600 python tools/record_probe.py > gpurun_out/r02l_record.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02l_record.txt
