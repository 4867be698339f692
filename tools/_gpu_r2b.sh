set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 600 python tools/sweep_r2.py --workloads c4-terrain,c4-soup,c2,c3 --tunes 0,0x8000 --shares 1 > gpurun_out/r2b_sweep_full.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x8000 --shares 2,4,8 > gpurun_out/r2b_sweep_shares.txt 2>&1
cat gpurun_out/r2b_sweep_full.txt gpurun_out/r2b_sweep_shares.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2b_bench.err
