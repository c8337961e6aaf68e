mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -1 gpurun_out/final_ref.json
