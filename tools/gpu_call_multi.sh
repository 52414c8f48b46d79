#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus 8): host->device ceiling of the box for GPU subsets, bench.py at N = 8, 4, 2
# (torchrun, one rank per GPU, with the in-bench ceiling probe and the config-4 leg), the in-library multi-device entry.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2; } > gpurun_out/r2m_topo.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG"
if [ "$NG" -ge 8 ]; then SUBS="0 0,1 0,1,2,3 4,5,6,7 0,2,4,6 0,1,2,3,4,5,6,7"; elif [ "$NG" -ge 4 ]; then SUBS="0 0,1 0,1,2,3"; else SUBS="0 0,1"; fi
( timeout 300 python tools/h2d_ceiling.py --mb 2048 --reps 8 $SUBS; timeout 300 python tools/h2d_ceiling.py --bind --mb 2048 --reps 8 $SUBS ) > gpurun_out/r2m_h2d_ceiling.jsonl 2> gpurun_out/r2m_h2d_ceiling.err
cat gpurun_out/r2m_h2d_ceiling.jsonl | cut -c1-260
for N in $NG $((NG/2)) 2; do
  [ "$N" -ge 2 ] || continue
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29$((500+N)) bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
  grep '^{' gpurun_out/r2m_bench_n$N.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); e = d['e2e']; c = d.get('config4') or {}
    print('N', d['n_gpus'], 'value', d['value'], 'e2e', e['value'], 'ms', e['ms_per_step'], 'h2d', e['h2d_gbs'], 'ceiling', e['h2d_ceiling_gbs'], 'frac', e['frac_of_h2d_ceiling'], e['host_binding'])
    print('  config4', c.get('value'), c.get('ms_total'), (c.get('e2e') or {}).get('value'), c.get('error'))"
  tail -2 gpurun_out/r2m_bench_n$N.err
done
( time timeout 600 python tools/multi_bench.py --streams 65536 --seconds 20 ) > gpurun_out/r2m_multi.json 2> gpurun_out/r2m_multi.err
cat gpurun_out/r2m_multi.json | cut -c1-600; tail -2 gpurun_out/r2m_multi.err
( time timeout 600 python -m pytest tests/test_engine_gpu.py -q -k "multi_device" ) > gpurun_out/r2m_tests.txt 2>&1
tail -3 gpurun_out/r2m_tests.txt
