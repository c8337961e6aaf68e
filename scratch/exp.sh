show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['iters_per_sec'],1), 'it/s', round(d['ms_per_step']*1e3,1), 'us/iter  kernel', round(d['roofline']['kernel_ms']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"; }
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
B="python bench.py --steps 100 --warmup 10 --no-cpu"
$B 2>&1 | show default
mkdir -p gpurun_out; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 12 --warmup 3 --no-cpu > /dev/null 2>&1
