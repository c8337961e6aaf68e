mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --opt tc_debug=8 > /dev/null 2> gpurun_out/g_dbg.err; grep nmfb200 gpurun_out/g_dbg.err | tail -2
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 3 > gpurun_out/g_cfg3.json 2> gpurun_out/g_cfg3.err; python -c "
import json; d=json.loads(open('gpurun_out/g_cfg3.json').read().strip().splitlines()[-1]); print('cfg3', round(d['iters_per_sec'],1), 'it/s', d['ms_per_step'], d['config'])"
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 2 > gpurun_out/g_cfg4.json 2> gpurun_out/g_cfg4.err; python -c "
import json; d=json.loads(open('gpurun_out/g_cfg4.json').read().strip().splitlines()[-1]); print('cfg4', round(d['iters_per_sec'],1), 'it/s', d['ms_per_step'], d['config'], d['coordinate_updates'])"
