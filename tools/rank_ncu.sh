# ncu --set full of the select-path kernels of one c3 evaluation: bash tools/rank_ncu.sh <out name> [kernel regex]
out="$1"; k="${2:-hamming_select_rank}"
timeout 300 python tools/map_probe.py c3 2 > gpurun_out/${out}_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/$out python tools/map_probe.py c3 2 > gpurun_out/${out}.log 2>&1
ls -la gpurun_out/$out.ncu-rep
