mkdir -p gpurun_out
timeout 600 python tools/sweep_r2.py --workloads c4-terrain,c3,c2 --tunes 0,0x8000,0x8200,0x1000 --shares 1 > gpurun_out/r2e_sweep.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x8200 --shares 8 >> gpurun_out/r2e_sweep.txt 2>&1
cat gpurun_out/r2e_sweep.txt
