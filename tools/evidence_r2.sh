#!/bin/bash
# Round-2 evidence run (executed on the B200 box through gpurun): tests, bench (both arms), ncu launch list and --set full
# captures of the hot kernels, full-size config probes.  Everything lands in gpurun_out/; tools/ncu_summary.py condenses the
# reports into profiles/.  Every profiler pass runs only after the same command exited 0 without the profiler.
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
SECONDS=0; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall=${SECONDS}s"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config3 --strong-log2"
timeout 300 $B > gpurun_out/plain_$TAG.log 2>&1 && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_a.log 2>&1
  # adapt_kernel launches of that command: 3 warm-up + 2 timed PARITY, then 3 + 2 FAST, ... : -s 3 is a timed PARITY launch, -s 8 a timed FAST one
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_parity $B > gpurun_out/ncu_b.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 8 -c 1 -f -o gpurun_out/prof_${TAG}_fast $B > gpurun_out/ncu_c.log 2>&1
}
timeout 900 python tests/gpu_configs.py all > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; cat gpurun_out/configs_$TAG.jsonl
# config 3 at its full size (2^22 trajectories, 67 GB of rows): launches are hinit, (coherence,) adapt — the second adapt launch is PARITY timed
C3="python tests/gpu_configs.py 3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_ -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_config3 $C3 > gpurun_out/ncu_d.log 2>&1
C4="python tests/gpu_configs.py 4 2"
timeout 300 $C4 > gpurun_out/configs4_small_$TAG.jsonl 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_config4 $C4 > gpurun_out/ncu_e.log 2>&1
Q="python tests/gpu_quick.py"
timeout 300 $Q > gpurun_out/quick_$TAG.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:adapt_kernel_f32 -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_f32 $Q > gpurun_out/ncu_f.log 2>&1
timeout 600 python tests/gpu_dense_probe.py > gpurun_out/dense_probe_$TAG.jsonl 2>&1; cat gpurun_out/dense_probe_$TAG.jsonl
head -c 600 gpurun_out/bench_$TAG.json; echo; ls -la gpurun_out | tail -14
# ---- also run by hand for the round-2 record (profiles/README.md):
#   differential campaign:  python tests/gpu_fuzz.py 20000 61 ; python tests/gpu_fuzz.py 1500 62 nvrtc ; python tests/gpu_fuzz.py 3000 63 fast   -> profiles/r2_fuzz_campaign.txt
#   2-GPU bench (gpurun --gpus 2): python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 -> profiles/r2_bench_2gpu.json
#   column-permutation microbenchmark: nvcc -O3 -std=c++17 [-DPERM_BK=4096] -gencode arch=compute_100a,code=sm_100a tools/micro/perm_bench.cu -o tools/micro/perm_bench ; tools/micro/perm_bench 20 500 -> profiles/r2_microbench_column_permutation.txt
