#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2i_variants.txt
cat > /tmp/v.py <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import sameold_b200 as sb
from sameold_b200 import synth
out = {}
for ns, secs, kernels in ((4096, 60.0, (3,)), (65536, 5.0, ())):
    n = int(secs * 22050); stride = (n + 7) // 8 * 8
    buf = torch.empty((ns, stride), dtype=torch.int16, device="cuda")
    synth.generate_on_device(synth.plan_corpus(ns, 22050, secs), buf.data_ptr(), stride, n)
    off = np.arange(ns, dtype=np.uint64) * np.uint64(stride); ln = np.full(ns, n, np.uint32)
    rx = sb.SameReceiverBuilder.samedec(22050).build_batch(ns)
    for k in kernels:
        rx.set_option("kernel", k)
        best = 1e9
        for r in range(4):
            rx.reset(); rx.submit_device(buf.data_ptr(), ns * stride, off, ln); rx.sync(); ev, _ = rx.drain_raw(reuse=True)
            if r: best = min(best, rx.last_timing()[1])
        out[f"k{k}_{ns}x{secs:g}_ms"] = round(best, 3); out[f"k{k}_events"] = int(ev.size)
    fe = rx.frontend_probe(buf.data_ptr(), ns * stride, off, ln, reps=10)
    out[f"fe_{ns}x{secs:g}_ms"] = round(fe, 3); out[f"fe_{ns}x{secs:g}_gbs"] = round(6.0 * ns * n / fe / 1e6, 1)
    del rx, buf; torch.cuda.empty_cache()
print(json.dumps(out))
PY
for lib in libsame_b200.so libfe1024.so libfe2048.so libpk2.so libsame_b200.so libpk2.so; do
  echo "$lib: $(SAME_B200_LIB=$PWD/sameold_b200/_build/$lib timeout 300 python /tmp/v.py 2>&1 | tail -1)" | tee -a gpurun_out/r2i_variants.txt
done
