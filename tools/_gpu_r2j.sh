mkdir -p gpurun_out
timeout 600 python tools/sweep_r2.py --workloads c4-terrain,c4-soup,c3,c2 --tunes 0,0x10000000 --shares 1 > gpurun_out/r2j_sweep_split.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x10000000 --shares 8 >> gpurun_out/r2j_sweep_split.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c2-area --tunes 0,0x10000000 --shares 1 --spp 16 >> gpurun_out/r2j_sweep_split.txt 2>&1
cat gpurun_out/r2j_sweep_split.txt
YAHR_B200_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_area_lights.py -m gpu -x -q > gpurun_out/r2j_pytest_split.log 2>&1; tail -4 gpurun_out/r2j_pytest_split.log
