mkdir -p gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "it/s", round(d["iters_per_sec"],1), "ms/step", round(d["ms_per_step"],4), "kern_ms", round(d["roofline"]["kernel_ms"],4), "e2e it/s", round(d["e2e"]["iters_per_sec"],1))
except Exception as e:
    print(tag, "FAILED", e)
PY
}
for rep in 1 2; do
( cd scratch/old_tree && timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu > ../../gpurun_out/b_old_$rep.json 2> ../../gpurun_out/b_old_$rep.err ); summ old_$rep gpurun_out/b_old_$rep.json
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --opt tc_pdl=0 > gpurun_out/b_new0_$rep.json 2> gpurun_out/b_new0_$rep.err; summ new_pdl0_$rep gpurun_out/b_new0_$rep.json
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/b_new1_$rep.json 2> gpurun_out/b_new1_$rep.err; summ new_pdl1_$rep gpurun_out/b_new1_$rep.json
done
( cd scratch/old_tree && timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu --timeline > /dev/null 2> ../../gpurun_out/b_old_tl.err ); echo OLD timeline; grep nmfb200 gpurun_out/b_old_tl.err | tail -3
timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu --timeline --opt tc_pdl=0 > /dev/null 2> gpurun_out/b_new_tl.err; echo NEW timeline; grep nmfb200 gpurun_out/b_new_tl.err | tail -3
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --opt tc_debug=8 > /dev/null 2> gpurun_out/b_dbg.err; grep nmfb200 gpurun_out/b_dbg.err | tail -4
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,temperature.gpu,power.draw --format=csv
