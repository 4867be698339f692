mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_fused -s 1 -c 1 -f -o gpurun_out/r2s_fused python tools/sweep_r2.py --workloads c4-terrain --tunes 0x1000 --reps 1 > gpurun_out/r2s_ncu2.log 2>&1; tail -1 gpurun_out/r2s_ncu2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err; echo "bench rc=$?"
