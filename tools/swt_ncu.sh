# ncu --set full of one SWT case for two planner settings: bash tools/swt_ncu.sh "<case substring>" "<ENV=.. for A>" "<ENV=.. for B>"
case="$1"; a="$2"; b="$3"
env $a timeout 300 python tools/swt_probe.py 2 "$case" > gpurun_out/swt_ncu_a_plain.log 2>&1 || exit 1
env $a timeout 600 ncu --set full --clock-control none --import-source on -k regex:swt2 --launch-skip 3 -c 1 -f -o gpurun_out/swt_ncu_a python tools/swt_probe.py 2 "$case" > gpurun_out/swt_ncu_a.log 2>&1
env $b timeout 600 ncu --set full --clock-control none --import-source on -k regex:swt2 --launch-skip 3 -c 1 -f -o gpurun_out/swt_ncu_b python tools/swt_probe.py 2 "$case" > gpurun_out/swt_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
