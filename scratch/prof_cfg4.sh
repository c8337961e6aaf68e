mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -s 20 -c 26 --csv --log-file gpurun_out/launches_cfg4_r1h.csv python bench.py --workload cfg4 --steps 3 --warmup 2 > /dev/null 2>&1
ls -la gpurun_out/launches_cfg4_r1h.csv
