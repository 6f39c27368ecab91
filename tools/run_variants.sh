#!/bin/bash
# On the GPU box: for each variants/*.so, install it as the library, run the bench, print kernel times.
cd "$(dirname "$0")/.."
cp openbts_ttsou_b200/libbtsdsp.so /tmp/libbtsdsp.keep
for v in "$@"; do
  [ "$v" != base ] && cp variants/$v.so openbts_ttsou_b200/libbtsdsp.so
  python bench.py --steps 10 --warmup 3 --no-e2e --no-secondary ${BENCH_ARGS:-} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k: round(x,4) for k, x in d['kernel_ms'].items()}, 'identical', (d['check']['vs_reference'] or {}).get('identical'))"
  cp /tmp/libbtsdsp.keep openbts_ttsou_b200/libbtsdsp.so
done
