mkdir -p gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "it/s", round(d["iters_per_sec"],1), "ms/step", round(d["ms_per_step"],4), "kern_ms", round(d["roofline"]["kernel_ms"],4), "launches", d["gpu_launches"])
except Exception as e:
    print(tag, "FAILED", e)
PY
}
( timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4 )
for fold in 0 1; do
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --opt tc_fold=$fold > gpurun_out/h_$fold.json 2> gpurun_out/h_$fold.err; summ fold$fold gpurun_out/h_$fold.json; tail -2 gpurun_out/h_$fold.err
done
