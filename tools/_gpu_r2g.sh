mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_smi.txt
run() { # name gpus extra-args env...
  name=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2g_bench_$name.json 2> gpurun_out/r2g_bench_$name.err
  echo "bench $name rc=$?"; tail -c 200 gpurun_out/r2g_bench_$name.err | grep -v resource_tracker | tail -2
}
run n8_auto 8
run n8_p2p 8 --exchange p2p --no-extra
YAHR_B200_HOST_FUSED=3 run n8_rows_persist 8 --exchange rows --no-extra
YAHR_B200_SHARD_STREAM=0 run n8_rows_nostream 8 --exchange rows --no-extra
run n4_auto 4
run n4_rows 4 --exchange rows --no-extra
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/dist_check.py > gpurun_out/r2g_dist_check_n8.log 2>&1; echo "dist_check n8 rc=$?"; grep -E "world=|OK" gpurun_out/r2g_dist_check_n8.log | tail -12
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_bench_*.json')):
    try: j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    x=j.get('extra_workloads') or [{}]
    print(f.split('/')[-1], 'value %.0f ms %.3f'%(j['value'],j['ms_per_step']), 'exch',j['config']['exchange'],'e2e %.3f rgb8 %.3f'%(j['e2e']['ms_per_step'], j['e2e']['rgb8']['ms_per_step']), {k:v for k,v in j['frame_check'].items() if 'single' in k}, 'c5', x[0].get('ms_per_step'), x[0].get('value'))
P
