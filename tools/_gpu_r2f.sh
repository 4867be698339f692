mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py > gpurun_out/r2f_dist_check_n2.log 2>&1; echo "dist_check rc=$?"; grep -v Warning gpurun_out/r2f_dist_check_n2.log | tail -14
for mode in auto rows p2p; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --exchange $mode --no-extra > gpurun_out/r2f_bench_n2_$mode.json 2> gpurun_out/r2f_bench_n2_$mode.err; echo "bench n2 $mode rc=$?"; tail -c 400 gpurun_out/r2f_bench_n2_$mode.err
done
YAHR_B200_SHARD_STREAM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --exchange rows --no-extra > gpurun_out/r2f_bench_n2_rows_nostream.json 2> gpurun_out/r2f_bench_n2_rows_nostream.err
timeout 600 python tools/sweep_r2.py --workloads c2-area --tunes 0 --shares 1 --spp 16 > gpurun_out/r2f_two_slot.txt 2>&1
YAHR_B200_TWO_SLOT=0 timeout 600 python tools/sweep_r2.py --workloads c2-area --tunes 0 --shares 1 --spp 16 >> gpurun_out/r2f_two_slot.txt 2>&1
timeout 600 python tools/sweep_r2.py --workloads c5-area --tunes 0 --shares 1 --spp 16 --reps 2 >> gpurun_out/r2f_two_slot.txt 2>&1
YAHR_B200_TWO_SLOT=0 timeout 600 python tools/sweep_r2.py --workloads c5-area --tunes 0 --shares 1 --spp 16 --reps 2 >> gpurun_out/r2f_two_slot.txt 2>&1
cat gpurun_out/r2f_two_slot.txt
