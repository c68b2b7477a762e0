T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 30 --warmup 5"
$T > gpurun_out/r02z_n8.json 2> gpurun_out/r02z_n8.err
tail -1 gpurun_out/r02z_n8.json | timeout 20 python tools/benchsum.py | cut -c1-90
VPD_NUMA_BIND=0 $T --no-configs > gpurun_out/r02z_n8_nobind.json 2>> gpurun_out/r02z_n8.err
tail -1 gpurun_out/r02z_n8_nobind.json | timeout 20 python tools/benchsum.py | cut -c1-90
python - <<'PY'
import json
for f in ['r02z_n8','r02z_n8_nobind']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['u8_batches']['value'], d.get('host_numa_bind'), d['dp_check'] and d['dp_check']['max_rel'], (d.get('configs') or {}).get('corpus_apply',{}).get('value'))
PY
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" 
