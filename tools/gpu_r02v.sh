# pose-encoder timeline at the headline shape (stamped twin of the library) + sampler parity
mkdir -p gpurun_out
timeout 300 python tools/diag_pose_timeline2.py > gpurun_out/r02v_pose_timeline.txt 2>&1; echo "rc=$?"; cat gpurun_out/r02v_pose_timeline.txt | cut -c1-700
timeout 900 python -m pytest tests/test_sampler.py tests/test_headline_parity.py -m gpu -x -q 2>&1 | tail -5
