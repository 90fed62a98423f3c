#!/usr/bin/env bash
# The standard on-box sequence of a measurement round (run through gpurun from the repo root; ALWAYS give gpurun a
# --timeout well below the budget that is left - a hung multi-GPU job is charged until the limit):
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r3a'           # tests, smoke, bench (SYN x 4096), reference arm, ncu
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r3a covid'     # + the cell kernel on COVID_C x 4096
# Everything lands in gpurun_out/ under the given tag; summarise the .ncu-rep files here with
#   python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep profiles/<tag>_grp_kernel_spec_fp32_syn_4096
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $out/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -1 $out/bench_$tag.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_${tag}_ref.log 2>&1
# launch list and one full capture of the dominant kernel (only after the plain run above exited 0); the env-group
# kernel takes ~1.7 s per launch, ~70 s under ncu --set full
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --no-secondary --cpu-sample-envs 1 --cpu-sample-steps 1 > $out/ncu_${tag}_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grp_kernel -s 3 -c 1 -f -o $out/prof_$tag \
    python bench.py --steps 4 --warmup 3 --no-secondary --cpu-sample-envs 1 --cpu-sample-steps 1 > $out/ncu_$tag.log 2>&1
if [ "${2:-}" = "covid" ]; then
  timeout 300 python bench.py --workload covid_c --steps 20 --warmup 5 --no-secondary > $out/bench_${tag}_covid.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_kernel -s 6 -c 1 -f -o $out/prof_${tag}_covid \
      python bench.py --workload covid_c --steps 10 --warmup 3 --no-secondary > $out/ncu_${tag}_covid.log 2>&1
fi
ls -la $out/*$tag* | tail -12
