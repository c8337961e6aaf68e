mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mu_update -s 8 -c 2 -o gpurun_out/prof_update_r1g python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/prof_r1g.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/*r1g*
