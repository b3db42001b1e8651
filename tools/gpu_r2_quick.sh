#!/bin/bash
# Round-2 iteration run (through gpurun): GPU parity tests, headline kernel timings, one ncu capture of the parity and the
# fast kernel for the instruction mix.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tests/gpu_quick.py > gpurun_out/quick_$TAG.log 2>&1; echo "quick rc=$?"; cat gpurun_out/quick_$TAG.log
Q="python tests/gpu_quick.py"
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__sass_inst_executed_op_global_ld.sum,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:adapt_kernel -c 12 --csv --log-file gpurun_out/ncu_mix_$TAG.csv $Q > gpurun_out/ncu_mix_$TAG.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,collections
rows=list(csv.reader(l for l in open("gpurun_out/ncu_mix_$TAG.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:90]), {})[r[mi]]=r[vi]
for k,v in d.items():
    try:
        wi=float(v["smsp__inst_executed.sum"].replace(",","")); f=float(v["sm__inst_executed_pipe_fp64.sum"].replace(",","")); ti=float(v["smsp__thread_inst_executed.sum"].replace(",",""))
        print(k[0], k[1][-60:], "ms=%.3f warp_inst=%.4g fp64=%.4g (%.1f%%) lanes=%.2f fp64pipe%%=%s regs=%s" % (float(v["gpu__time_duration.sum"].replace(",",""))/1e6, wi, f, 100*f/wi, ti/wi, v["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"], v["launch__registers_per_thread"]))
    except Exception as e:
        print(k, v, e)
PY
