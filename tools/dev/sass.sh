#!/bin/bash
# usage: sass.sh parity|fast [extra nvcc flags]  -> /tmp/sass/<name>.sass + opcode histogram
set -e
name=$1; shift
cd "$(dirname "$0")"
if [ "$name" = fast ]; then F="-fmad=true -DDEQ_FAST_TU -DHOT_FAST=true"; else F="-fmad=false"; fi
/usr/local/cuda/bin/nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 $F --expt-relaxed-constexpr -Xptxas -v "$@" -cubin hot.cu -o /tmp/sass/$name.cubin 2>&1 | grep -E "registers|spill|error" || true
cuobjdump -sass /tmp/sass/$name.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/sass/$name.sass
wc -l /tmp/sass/$name.sass
