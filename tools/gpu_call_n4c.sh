#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/h2d_ceiling.py --rows 4096 --width 110250 --pitch 2646016 --reps 48 0,4,1,5 0,1,2,3 4,5,6,7 > gpurun_out/r2n4c_h2d.jsonl 2>/dev/null
cut -c1-330 gpurun_out/r2n4c_h2d.jsonl
for N in 4 2; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29$((700+N)) bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2n4c_bench_n$N.json 2> gpurun_out/r2n4c_bench_n$N.err
  grep '^{' gpurun_out/r2n4c_bench_n$N.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); e = d['e2e']; c = d.get('config4') or {}
    print('N', d['n_gpus'], 'value', d['value'], 'e2e', e['value'], 'ms', e['ms_per_step'], 'h2d', e['h2d_gbs'], 'ceiling', e['h2d_ceiling_gbs'], 'frac', e['frac_of_h2d_ceiling'], e['host_binding'])
    print('  config4', c.get('value'), c.get('ms_total'), (c.get('e2e') or {}).get('value'), c.get('error'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
  tail -2 gpurun_out/r2n4c_bench_n$N.err
done
