set -u
O=gpurun_out/r02g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gputest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"; tail -2 $O/bench_1gpu.err | cut -c1-300
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python tools/one_step.py > $O/one_step_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/one_step.py > $O/ncu_step.log 2>&1
echo "launch list rc=$?"; tail -1 $O/one_step_plain.log
timeout 300 python tools/one_step.py --greedy > $O/one_greedy_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_greedy.csv python tools/one_step.py --greedy > $O/ncu_greedy.log 2>&1
echo "greedy launch list rc=$?"; tail -1 $O/one_greedy_plain.log
timeout 300 python tools/one_step.py > $O/one_step_plain2.log 2>&1 &&
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $O/step_full python tools/one_step.py > $O/ncu_full.log 2>&1
echo "full rc=$?"
ncu -i $O/step_full.ncu-rep --page raw --csv > $O/step_full_raw.csv 2>/dev/null
timeout 300 python tools/one_step.py --greedy > $O/one_greedy_plain2.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'gemm_tc_kernel|argmax_finish' -s 30 -c 6 -f -o $O/greedy_full python tools/one_step.py --greedy > $O/ncu_greedy_full.log 2>&1
ncu -i $O/greedy_full.ncu-rep --page raw --csv > $O/greedy_full_raw.csv 2>/dev/null
rm -f $O/step_full.ncu-rep $O/greedy_full.ncu-rep
ls -la $O | tail -20
