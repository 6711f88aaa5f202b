set -u
O=gpurun_out/r02g; mkdir -p $O
for v in default noprep prepmain; do
  case $v in default) E="";; noprep) E="SNT_NO_EARLY_PREP=1";; prepmain) E="SNT_PREP_MAIN=1";; esac
  echo "== $v"; env $E timeout 200 python tools/e2e_probe.py 40 2>&1 | grep -v "per-call (ms)" | tail -6
done
