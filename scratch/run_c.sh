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
( cd scratch/old_tree && timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu > ../../gpurun_out/c_old.json 2> ../../gpurun_out/c_old.err ); summ old gpurun_out/c_old.json
for dbg in 0 16 32 64 128 240 48 208; do
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --opt tc_pdl=0 --opt tc_debug=$dbg > gpurun_out/c_$dbg.json 2> gpurun_out/c_$dbg.err; summ pdl0_dbg$dbg gpurun_out/c_$dbg.json
done
for dbg in 0 32 64 96; do
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --opt tc_pdl=1 --opt tc_debug=$dbg > gpurun_out/c1_$dbg.json 2> gpurun_out/c1_$dbg.err; summ pdl1_dbg$dbg gpurun_out/c1_$dbg.json
done
