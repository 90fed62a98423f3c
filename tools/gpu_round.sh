#!/usr/bin/env bash
# The standard on-box sequence of a measurement round (run through gpurun from the repo root):
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh r2a'          # tests, smoke, bench, reference arm, ncu captures
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh r2a syn'      # + the env kernel on SYN (slow under ncu: ~10 min)
# Everything lands in gpurun_out/ under the given tag; summarise the .ncu-rep files here with
#   python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep profiles/<tag>_cell_kernel_spec_fp32_covid_c_4096
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
python bench.py > $out/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -1 $out/bench_$tag.log | cut -c1-200
python bench.py --impl reference --steps 10 --warmup 3 > $out/bench_${tag}_ref.log 2>&1
# launch list and one full capture of the dominant kernel (only after the plain run above exited 0)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 10 --warmup 3 > $out/ncu_${tag}_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_kernel -s 6 -c 1 -f -o $out/prof_$tag \
    python bench.py --steps 10 --warmup 3 > $out/ncu_$tag.log 2>&1
if [ "${2:-}" = "syn" ]; then
  python bench.py --workload syn --envs 592 --steps 3 --warmup 3 --cpu-sample-envs 1 --cpu-sample-steps 1 > $out/bench_${tag}_syn592.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 3 -c 1 -f -o $out/prof_${tag}_syn \
      python bench.py --workload syn --envs 592 --steps 3 --warmup 3 --cpu-sample-envs 1 --cpu-sample-steps 1 > $out/ncu_${tag}_syn.log 2>&1
fi
ls -la $out/*$tag* | tail -12
