mkdir -p gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "it/s", round(d["iters_per_sec"],1), "ms/step", round(d["ms_per_step"],4), "kern_ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e:
    print(tag, "FAILED", e)
PY
}
( timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3 )
for rep in 1 2; do for pdl in 0 1; do
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --opt tc_pdl=$pdl > gpurun_out/e_$pdl.json 2> gpurun_out/e_$pdl.err; summ pdl$pdl gpurun_out/e_$pdl.json
done; done
