# End-of-round ncu evidence (one GPU): launch list of the c3 bench step, --set full of the packing kernels and of the
# select pipeline.  Every ncu command follows a plain run of the same program that exited 0.
set -x
timeout 600 python bench.py --no-extras --steps 2 --warmup 1 > gpurun_out/r2z_bench_plain.json 2> gpurun_out/r2z_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_bench_c3.csv \
    python bench.py --no-extras --steps 2 --warmup 1 > gpurun_out/r2z_ncu_launches.log 2>&1
timeout 300 python tools/pack_probe.py > gpurun_out/r2z_pack_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pack_rows -c 8 -f -o gpurun_out/prof_r2z_pack \
    python tools/pack_probe.py > gpurun_out/r2z_ncu_pack.log 2>&1
timeout 300 python tools/map_probe.py c3 2 > gpurun_out/r2z_map_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_sample|select_bound|hamming_select" --launch-skip 4 -c 4 -f \
    -o gpurun_out/prof_r2z_select python tools/map_probe.py c3 2 > gpurun_out/r2z_ncu_select.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
