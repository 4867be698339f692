mkdir -p gpurun_out
echo "# row order ON (most expensive rows first, previous frame's costs)" > gpurun_out/r2l_row_order.txt
timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x1000 --shares 2,4,8 --reps 8 >> gpurun_out/r2l_row_order.txt 2>&1
echo "# row order OFF (YAHR_B200_NO_ROW_ORDER=1)" >> gpurun_out/r2l_row_order.txt
YAHR_B200_NO_ROW_ORDER=1 timeout 300 python tools/sweep_r2.py --workloads c4-terrain --tunes 0,0x1000 --shares 2,4,8 --reps 8 >> gpurun_out/r2l_row_order.txt 2>&1
cat gpurun_out/r2l_row_order.txt
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -m gpu -x -q -k "dist or streamed or host_entries" > gpurun_out/r2l_pytest.log 2>&1; tail -3 gpurun_out/r2l_pytest.log
