#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, then (only if those exit 0)
# the ncu launch list and one full capture of the top kernel.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
T=$? ; tail -n 25 gpurun_out/pytest_gpu.log ; echo "pytest rc=$T"
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
S=$? ; tail -n 5 gpurun_out/smoke.log ; echo "smoke rc=$S"
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-30} --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
B=$? ; tail -n 3 gpurun_out/bench.log ; tail -n 5 gpurun_out/bench.err ; echo "bench rc=$B"
if [ "${RUN_NCU:-1}" = "1" ] && [ $T -eq 0 ] && [ $B -eq 0 ]; then
  echo "== ncu launch list"
  timeout 600 python bench.py --steps 3 --warmup 3 --no_eval --no_cpu_baseline > gpurun_out/plain_small.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no_eval --no_cpu_baseline \
      > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  if [ -n "${NCU_KERNEL:-}" ]; then
    echo "== ncu full: $NCU_KERNEL"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -s 3 -c 2 \
        -f -o gpurun_out/prof python bench.py --steps 3 --warmup 3 --no_eval --no_cpu_baseline \
        > gpurun_out/ncu_full.log 2>&1
    echo "ncu full rc=$?"
  fi
fi
exit 0
