N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py 2>&1 | grep -E "rank 0.*tc|dist_gpu_check ok|FAIL|rror|timed out" | tail -6
for pdl in 1 0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --opt tc_pdl=$pdl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N pdl=$pdl', round(d['iters_per_sec'],1), 'it/s', round(d['ms_per_step']*1e3,1), 'us/iter', 'e2e', round(d['e2e']['seconds'],4), 'objv', d['config']['objvalue_e2e'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --no-e2e --timeline 2>&1 | grep -E "rank 0\] phase"
