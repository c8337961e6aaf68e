# r1g: PDL (programmatic dependent launch) update kernels + denominator-last block order
mkdir -p gpurun_out
( timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/a_tests.log 2>&1
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/a_bench_pdl1.json 2> gpurun_out/a_bench_pdl1.err
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --opt tc_pdl=0 > gpurun_out/a_bench_pdl0.json 2> gpurun_out/a_bench_pdl0.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --opt tc_debug=8 > gpurun_out/a_bench_dbg.json 2> gpurun_out/a_bench_dbg.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --opt tc_debug=8 --opt tc_pdl=0 > gpurun_out/a_bench_dbg0.json 2> gpurun_out/a_bench_dbg0.err
cat gpurun_out/a_tests.log
for f in pdl1 pdl0 dbg dbg0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/a_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "it/s", round(d["iters_per_sec"],1), "ms/step", round(d["ms_per_step"],4), "kern_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "e2e it/s", round(d["e2e"]["iters_per_sec"],1), "objv", d["config"]["objvalue_e2e"], d["clocks"])
except Exception as e:
    print("$f", "FAILED", e)
PY
tail -3 gpurun_out/a_bench_$f.err
done
