cd /root/repo
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err
grep '^{' gpurun_out/r2z_bench_n2.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); e = d['e2e']; c = d.get('config4') or {}
    print('N', d['n_gpus'], 'value', d['value'], 'e2e', e['value'], 'ms', e['ms_per_step'], 'h2d', e['h2d_gbs'], 'ceiling', e['h2d_ceiling_gbs'], e['host_binding'])
    print('  config4', c.get('value'), c.get('ms_total'), (c.get('e2e') or {}).get('value'), c.get('error'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
tail -3 gpurun_out/r2z_bench_n2.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r2z_ref_n2.json 2> gpurun_out/r2z_ref_n2.err; tail -c 300 gpurun_out/r2z_ref_n2.json
