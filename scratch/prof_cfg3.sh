mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 24 --csv --log-file gpurun_out/launches_cfg3_r1h.csv python bench.py --workload cfg3 --steps 6 --warmup 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:div_fused -s 4 -c 2 -o gpurun_out/prof_divfused_r1h python bench.py --workload cfg3 --steps 4 --warmup 2 > gpurun_out/prof_divfused.log 2>&1
ls -la gpurun_out/*r1h*
