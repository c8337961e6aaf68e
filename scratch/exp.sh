show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['iters_per_sec'],1), 'it/s', round(d['ms_per_step']*1e3,1), 'us/iter  kernel', round(d['roofline']['kernel_ms']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"; }
B="python bench.py --steps 100 --warmup 10 --no-cpu"
for pf in 0 8 16 32 64; do $B --opt tc_prefetch=$pf 2>&1 | show prefetch$pf; done
