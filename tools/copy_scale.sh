#!/bin/bash
# copy-only ceiling of the box at N = 1, 2, 4, 8 concurrent GPUs (tools/copy_probe, one process per GPU)
# usage: tools/copy_scale.sh <outdir> [h2d_MB d2h_MB]
out=${1:-gpurun_out/copy_scale}; h2d=${2:-739}; d2h=${3:-131}
mkdir -p "$out"
ngpu=$(nvidia-smi -L | wc -l)
(lscpu | head -30; nvidia-smi topo -m; ls /sys/devices/system/node/; free -g) > "$out/topo.txt" 2>&1
for n in 1 2 4 8; do
  [ "$n" -gt "$ngpu" ] && break
  for ((g = 0; g < n; g++)); do tools/copy_probe $g $h2d $d2h 5 > "$out/n${n}_gpu${g}.json" 2>&1 & done
  wait
done
