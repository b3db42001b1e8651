#!/bin/bash
# Round-2 evidence run (executed on the B200 box through gpurun): tests, bench (both arms), ncu launch list and --set full
# captures of the hot kernels.  Everything lands in gpurun_out/; tools/ncu_summary.py condenses the reports into profiles/.
# Every profiler pass runs only after the same command exited 0 without the profiler.
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
SECONDS=0; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall=${SECONDS}s"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config3 --strong-log2"
timeout 300 $B > gpurun_out/plain_$TAG.log 2>&1 && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_a.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_parity $B > gpurun_out/ncu_b.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 8 -c 1 -f -o gpurun_out/prof_${TAG}_fast $B > gpurun_out/ncu_c.log 2>&1
}
head -c 1500 gpurun_out/bench_$TAG.json; echo; ls -la gpurun_out | tail -12
