#!/bin/bash
# tile-shape sweep for the SWT planner (B200_SWT_TILE override), prints CUDA-event times per case
for tile in "" "16,112" "32,112" "16,224" "32,56" "32,64" "16,260" "32,104" "48,104" "48,76" "64,64" "32,88" "16,132" "32,132"; do
  echo "== tile [$tile] threads [${B200_SWT_THREADS}]"
  B200_SWT_TILE=$tile python tools/swt_cases.py 3 ${1:-0,2,3,4,5,6}
done
