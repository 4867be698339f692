mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compressed or small_scenes or soup or degenerate" > gpurun_out/r2i_pytest.log 2>&1; tail -15 gpurun_out/r2i_pytest.log
timeout 600 python tools/sweep_r2.py --workloads c4-soup,c4-terrain,c3,c2 --tunes 0,0x40000000 --shares 1 > gpurun_out/r2i_sweep_compressed.txt 2>&1; cat gpurun_out/r2i_sweep_compressed.txt
