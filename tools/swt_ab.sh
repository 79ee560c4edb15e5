# A/B of the SWT planner knobs over bench.py's C4 grid: bash tools/swt_ab.sh
run() { echo "== $*"; env "$@" timeout 300 python tools/swt_probe.py 10 "$CASE" 2>&1 | grep -v "^\*\|OMP"; }
CASE="L2"
run B200_SWT_VS=1
for t in 32,76 64,76 32,104 48,104; do run B200_SWT_VS=1 B200_SWT_TILE=$t; done
CASE="L3"
for t in 32,76 64,76 32,104 64,104; do run B200_SWT_VS=1 B200_SWT_TILE=$t; done
