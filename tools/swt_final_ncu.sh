# ncu --set full of one launch of the final SWT kernels on cases of the C4 grid (after a plain run that exited 0); the
# reports are summarised on the box (tools/ncu_summary.py) and removed: six of them exceed what gpurun copies back.
timeout 600 python tools/swt_probe.py 3 > gpurun_out/r2z_swt_plain.log 2>&1 || exit 1
i=0
for c in "haar L1 518" "db4 L1 518" "db4 L2" "sym4 L3"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:swt2 --launch-skip 3 -c 1 -f -o /tmp/prof_r2z_swt_$i \
      python tools/swt_probe.py 2 "$c" > gpurun_out/r2z_swt_ncu_$i.log 2>&1
  python tools/ncu_summary.py rep /tmp/prof_r2z_swt_$i.ncu-rep gpurun_out/r2z_swt_$i.md "$c" > /dev/null 2>&1
  i=$((i+1))
done
ls -la gpurun_out/r2z_swt_*.md
