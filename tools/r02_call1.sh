set -u
mkdir -p gpurun_out/r02
bash tools/switch_sweep.sh > gpurun_out/r02/sweep_summary.txt 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_persistent.py > gpurun_out/r02/sanitize_${tool}_h256.log 2>&1
  echo "$tool h256 rc=$?" >> gpurun_out/r02/sweep_summary.txt
done
SNT_SAN_H=512 SNT_SAN_B=300 timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_persistent.py > gpurun_out/r02/sanitize_racecheck_h512.log 2>&1
echo "racecheck h512 rc=$?" >> gpurun_out/r02/sweep_summary.txt
cat gpurun_out/r02/sweep_summary.txt
