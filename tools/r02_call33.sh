set -u
O=gpurun_out/r02l; mkdir -p $O
for v in default nosplit; do
  case $v in default) E="SNT_PREP_DEBUG=1";; nosplit) E="SNT_PREP_DEBUG=1 SNT_NO_SPLIT_PROJECTION=1";; esac
  echo "== $v"; env $E timeout 200 python tools/e2e_probe.py 40 > $O/$v.log 2>&1
  grep "prep dbg" $O/$v.log | sed -n '2,5p;22,24p'; grep -v "per-call (ms)\|prep dbg" $O/$v.log | tail -5
done
