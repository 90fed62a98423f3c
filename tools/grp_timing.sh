#!/usr/bin/env bash
# Critical-path breakdown of the env-group kernel (debug build with -DEPI_GRP_TIMING): builds the timing variant of the
# SYN fp32 specialised library HERE, swaps it in, runs a small bench through gpurun, restores the regular build.
#   tools/grp_timing.sh [extra env, e.g. EPI_GRP_WARPS=12]
set -e
cd "$(dirname "$0")/.."
SO=$(env $1 python - <<'PY'
import sys; sys.path.insert(0,'.')
from epipolicy_rl_b200 import spec, scenarios
d=scenarios.load('syn'); text,key=spec.source(d,1)
open('/tmp/spec_syn_t.h','w').write(text); print(spec.so_path(key))
PY
)
cp "$SO" /tmp/spec_syn_regular.so
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -w -DEPI_GRP_TIMING '-DEPI_SPEC_HEADER="/tmp/spec_syn_t.h"' -I epipolicy_rl_b200/csrc -o "$SO" epipolicy_rl_b200/csrc/epi_spec_main.cu
gpurun --timeout 600 -- "env $1 timeout 300 python bench.py --envs 64 --steps 2 --warmup 3 --no-secondary --cpu-sample-envs 1 --cpu-sample-steps 1 2>&1 | grep 'grp timing' | tail -1" 2>&1 | grep "grp timing\|GPU-minutes" | tee /tmp/grp_timing.out
cp /tmp/spec_syn_regular.so "$SO"
python3 - <<'PY'
import re
names={0:'other',1:'pass1',2:'arrive',3:'flows_reg',4:'waitA',5:'pass2 gather',6:'pass3 mixing',7:'barrier B',8:'pass4a gather+inf',9:'pass4b',10:'grp_sum',11:'norms',12:'commit',13:'fill_tables(dyn)',14:'prologue'}
t=open('/tmp/grp_timing.out').read()
v={}
for m in re.finditer(r's(\d+)=([0-9.]+)ms', t):
    v.setdefault(int(m.group(1)),[]).append(float(m.group(2)))
tot=sum(sum(x)/len(x) for x in v.values())
print(f"total {tot:.2f} ms per launch (16 env-steps per group) = {tot/16*1e3:.0f} us per env-step")
for k in sorted(v): 
    a=sum(v[k])/len(v[k]); print(f"  {names.get(k,k):20s} {a/16*1e3:7.1f} us/env-step  {100*a/tot:5.1f}%   {a/16*1e3/92:5.2f} us/RHS")
PY
