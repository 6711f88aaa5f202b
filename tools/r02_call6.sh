set -u
O=gpurun_out/r02
mkdir -p $O
timeout 300 python tools/e2e_probe.py 40 > $O/e2e_probe.txt 2>&1; echo "e2e_probe rc=$?"; cat $O/e2e_probe.txt | cut -c1-400
timeout 300 python tools/one_step.py > $O/one_step_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/one_step.py > $O/ncu_step.log 2>&1
echo "launch list rc=$?"; cat $O/one_step_plain.log | tail -2
timeout 300 python tools/one_step.py --greedy > $O/one_greedy_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_greedy.csv python tools/one_step.py --greedy > $O/ncu_greedy.log 2>&1
echo "greedy launch list rc=$?"; tail -1 $O/one_greedy_plain.log
timeout 300 python tools/one_step.py > $O/one_step_plain2.log 2>&1 &&
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $O/step_full python tools/one_step.py > $O/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 $O/ncu_full.log
ncu -i $O/step_full.ncu-rep --page raw --csv > $O/step_full_raw.csv 2>/dev/null; ls -la $O
sz=$(stat -c %s $O/step_full.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 45000000 ]; then rm -f $O/step_full.ncu-rep; echo "rep dropped ($sz bytes)"; fi
