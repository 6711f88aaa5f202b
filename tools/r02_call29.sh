set -u
O=gpurun_out/r02h; mkdir -p $O
for v in default prio; do
  case $v in default) E="SNT_PREP_DEBUG=1";; prio) E="SNT_PREP_DEBUG=1 SNT_SIDE2_PRIO=-2";; esac
  echo "== $v"; env $E timeout 200 python tools/e2e_probe.py 40 > $O/$v.log 2>&1
  grep "prep dbg" $O/$v.log | sed -n '1,3p;8,10p;14,16p;22,24p'; grep -v "per-call (ms)\|prep dbg" $O/$v.log | tail -5
done
