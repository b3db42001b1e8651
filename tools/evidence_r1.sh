#!/bin/bash
# Round-1 evidence run (executed on the B200 box through gpurun): tests, bench (both arms), ncu launch list and --set full
# captures of the hot kernels, full-size config probes.  Everything lands in gpurun_out/; tools/ncu_summary.py condenses the
# reports into profiles/.  Every profiler pass runs only after the same command exited 0 without the profiler.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain.log 2>&1 && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_a.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_parity_final $B > gpurun_out/ncu_b.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 8 -c 1 -f -o gpurun_out/prof_r1_fast_final $B > gpurun_out/ncu_c.log 2>&1
}
S="python tests/gpu_stiff_bench.py 1048576"
timeout 300 $S > gpurun_out/stiff_probe.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ode23s_kernel -s 5 -c 1 -f -o gpurun_out/prof_r1_ode23s_textbook $S > gpurun_out/ncu_d.log 2>&1
timeout 900 python tests/gpu_configs.py all > gpurun_out/configs_r1.jsonl 2> gpurun_out/configs_r1.err
timeout 600 python tests/gpu_dense_unsorted.py > gpurun_out/dense_unsorted_r1.jsonl 2> gpurun_out/dense_unsorted_r1.err
tail -3 gpurun_out/pytest_gpu.log; head -c 600 gpurun_out/bench_r1_final.json; echo; cat gpurun_out/stiff_probe.log | tail -8; ls -la gpurun_out | tail -20
