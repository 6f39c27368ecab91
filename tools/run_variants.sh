#!/bin/bash
# On the GPU box: for each variants/*.so, install it as the library, run the bench, print kernel times.
cd "$(dirname "$0")/.."
cp openbts_ttsou_b200/libbtsdsp.so /tmp/libbtsdsp.keep
for v in "$@"; do
  [ "$v" != base ] && cp variants/$v.so openbts_ttsou_b200/libbtsdsp.so
  python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), [(k['name'][:12],round(k['ms'],4)) for k in d['roofline']['kernels'][:3]], d['check'], 'e2e_wire', round(d['e2e_wire']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
  cp /tmp/libbtsdsp.keep openbts_ttsou_b200/libbtsdsp.so
done
