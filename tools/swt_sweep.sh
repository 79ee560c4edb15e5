#!/bin/bash
# tile-shape sweep for the SWT planner (B200_SWT_TILE override), prints CUDA-event times per case
for tile in "" "16,112" "32,112" "16,224" "32,56" "32,128" "48,128" "16,520" "16,260" "32,260" "24,260" "32,176" "32,104" "48,104" "48,76" "64,64" "32,64" "96,40" "64,88" "32,88"; do
  echo "== tile [$tile] threads [${B200_SWT_THREADS}]"
  B200_SWT_TILE=$tile python tools/swt_cases.py 3 ${1:-0,2,3,4,5,6}
done
