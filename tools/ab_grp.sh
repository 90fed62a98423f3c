#!/usr/bin/env bash
# A/B of the env-group kernel's layout variants on SYN (run through gpurun): tools/ab_grp.sh <tag> [steps] [cfg;cfg;...]
tag=${1:-rX}; steps=${2:-4}; cfgs=${3:-"EPI_AB=default;EPI_GRP_WARPS=12"}
mkdir -p gpurun_out
i=0
IFS=';' read -ra CF <<< "$cfgs"
for cfg in "${CF[@]}"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --steps $steps --warmup 3 --no-secondary --cpu-sample-envs 1 --cpu-sample-steps 1 > gpurun_out/ab_${tag}_$i.log 2>&1
  echo "== $cfg rc=$?"
  tail -1 gpurun_out/ab_${tag}_$i.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"launch": {[^}]*}' | head -3 | tr '\n' ' '; echo
done
