set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python tools/sweep_r2.py --workloads c4-terrain,c4-soup,c2,c3 --tunes 0,0x8000000,0xC000000,0x8000,0x8008000,0xC008000,0x1C008000,0x400 --shares 1 > gpurun_out/r2a_sweep_full.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0xC000000,0x8000,0xC008000 --shares 2,4,8 > gpurun_out/r2a_sweep_shares.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.err
cat gpurun_out/r2a_sweep_full.txt
