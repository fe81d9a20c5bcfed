set -x
mkdir -p gpurun_out
timeout 600 python tools/jitter_probe.py 40 > gpurun_out/r02i_jitter.txt 2>&1; echo "probe rc=$?"; cut -c1-700 gpurun_out/r02i_jitter.txt
