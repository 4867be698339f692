# Final multi-GPU evidence of a round on an 8-GPU box: bench at N = 8, 4, 2 (default exchange = p2p; the row pushes at 8 too), each line
# with frame_check and extra_workloads, and the stand-alone frame-assembly check at N = 8.  Results in gpurun_out/ (r2y_*).
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2y_bench_$name.json 2> gpurun_out/r2y_bench_$name.err
  echo "bench $name rc=$?"; }
run n8 8
run n8_rows 8 --exchange rows --no-extra
run n4 4
run n2 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tests/dist_check.py > gpurun_out/r2y_dist_check_n8.log 2>&1; echo "dist_check n8 rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y_bench_*.json')):
    try: j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value %.0f ms %.3f'%(j['value'],j['ms_per_step']), 'exch',j['config']['exchange'],'e2e %.3f rgb8 %.3f'%(j['e2e']['ms_per_step'], j['e2e']['rgb8']['ms_per_step']), {k:v for k,v in j['frame_check'].items() if 'single' in k}, [(x['name'],round(x['ms_per_step'],2)) for x in (j.get('extra_workloads') or [])])
P
