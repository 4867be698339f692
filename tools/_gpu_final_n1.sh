# Final single-GPU evidence of a round: tests, smoke, bench (twice: plain and the one the launch list is taken from), the
# ncu launch list and one `ncu --set full` capture of the default kernels.  Results land in gpurun_out/ (r2z_*).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -3 gpurun_out/r2z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2z_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 6 -c 3 -o gpurun_out/r2z_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-counts > gpurun_out/r2z_ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/r2z_full.ncu-rep
