mkdir -p gpurun_out
for tr in 128 112; do
timeout 300 ncu --set full --clock-control none -k regex:mu_update -s 8 -c 1 -o gpurun_out/prof_tr$tr python bench.py --steps 6 --warmup 3 --no-cpu --opt tc_tile_rows=$tr > gpurun_out/prof_tr$tr.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
