set -x
mkdir -p gpurun_out
timeout 600 python tools/sweep_r2.py --workloads c4-terrain,c4-soup,c2,c3 --tunes 0,0x8000,0x1000 --shares 1 > gpurun_out/r2c_sweep_full.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x8000 --shares 2,4,8 > gpurun_out/r2c_sweep_shares.txt 2>&1
cat gpurun_out/r2c_sweep_full.txt gpurun_out/r2c_sweep_shares.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "streamed or host_entries or rgb8 or dist" > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
