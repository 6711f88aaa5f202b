set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s > $O/gputest_i.log 2>&1; echo "pytest fullsize rc=$?"; grep "rel err\|passed\|failed\|Error" $O/gputest_i.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest_j.log 2>&1; echo "pytest rc=$?"; tail -4 $O/gputest_j.log
