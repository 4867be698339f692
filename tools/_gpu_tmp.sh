mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
timeout 300 python tools/sweep_r2.py --workloads c1,c2,c3,c4-terrain --tunes 0,0x20000000,0x1000 --shares 1 --reps 6 > gpurun_out/r2q_auto.txt 2>&1
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x20000000 --shares 2,4,8 --reps 6 >> gpurun_out/r2q_auto.txt 2>&1
cat gpurun_out/r2q_auto.txt
