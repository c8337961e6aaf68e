mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "multdiv" 2>&1 | tail -15 )
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "config3" 2>&1 | tail -3 )
for f in 1 0; do
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 3 --opt tc_div_fused=$f > gpurun_out/i_cfg3_$f.json 2> gpurun_out/i_cfg3_$f.err; python -c "
import json; d=json.loads(open('gpurun_out/i_cfg3_$f.json').read().strip().splitlines()[-1]); print('cfg3 fused=$f', round(d['iters_per_sec'],1), 'it/s', round(d['ms_per_step'],4), 'objv', d['objvalue'], 'launches', d['gpu_launches'])"; tail -2 gpurun_out/i_cfg3_$f.err
done
