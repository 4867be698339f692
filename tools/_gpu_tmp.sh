mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err; echo "bench rc=$?"
python - <<'P'
import json
j=json.loads(open('gpurun_out/r2z2_bench.json').read().strip().splitlines()[-1])
r=j['roofline']
print(j['value'], j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e']['rgb8']['ms_per_step'])
print({k:r[k] for k in ('bound','achieved','peak','unit','frac','traffic','counts_source','counts_refused')})
print(j['frame_check'])
for x in j['extra_workloads']: print(x['name'],x['spp'],round(x['ms_per_step'],3),round(x['value']))
P
