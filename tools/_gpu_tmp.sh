mkdir -p gpurun_out
echo "# 8-bit rows from the two-kernel set (default)" > gpurun_out/r2o_rgb8_two_kernel.txt
timeout 300 python tools/e2e_held_sweep.py c4-terrain 2>&1 | grep -E "6144" >> gpurun_out/r2o_rgb8_two_kernel.txt
echo "# 8-bit rows from k_wf_fused (YAHR_B200_HOST_FUSED=1)" >> gpurun_out/r2o_rgb8_two_kernel.txt
YAHR_B200_HOST_FUSED=1 timeout 300 python tools/e2e_held_sweep.py c4-terrain 2>&1 | grep -E "6144" >> gpurun_out/r2o_rgb8_two_kernel.txt
for w in c2 c3; do
echo "# $w: two-kernel / fused" >> gpurun_out/r2o_rgb8_two_kernel.txt
timeout 300 python tools/e2e_held_sweep.py $w 2>&1 | grep -E "6144" >> gpurun_out/r2o_rgb8_two_kernel.txt
YAHR_B200_HOST_FUSED=1 timeout 300 python tools/e2e_held_sweep.py $w 2>&1 | grep -E "6144" >> gpurun_out/r2o_rgb8_two_kernel.txt
done
timeout 300 python tools/sweep_r2.py --workloads c4-terrain,c2 --tunes 0 --shares 1 >> gpurun_out/r2o_rgb8_two_kernel.txt 2>&1
cat gpurun_out/r2o_rgb8_two_kernel.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
