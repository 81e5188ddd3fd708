#!/bin/bash
# First GPU run of round 2, one gpurun call (everything below was written after round 1's GPU budget ran out):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_run.sh'
# Every step has its own timeout (the experimental kernels trap after ~2 s instead of hanging, but have never run) and
# writes into gpurun_out/r2_first/.  Nothing here changes defaults; read the outputs, then decide what to switch on.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
run() { # name, timeout, command...
    local name=$1 t=$2; shift 2
    echo "=== $name" | tee -a "$OUT/summary.txt"
    timeout "$t" "$@" > "$OUT/$name.log" 2>&1
    echo "rc=$? ($(tail -1 "$OUT/$name.log" | cut -c1-200))" | tee -a "$OUT/summary.txt"
}
# 1. gated silence-mask kernel: bit-exact vs the oracle?
STB_UNVERIFIED_KERNELS=1 run silence_tests 300 python -m pytest tests/test_gpu_silence.py -x -q
# 2. cross-attention V2 (half2 residual dot, block-wise rescale): parity, then speed
STB_XATTN_V2=1 run xattn_v2_tests 400 python -m pytest tests/test_gpu_decode.py -x -q -k "not above_64"
# 3. persistent chain kernel for the decode linears: parity on the tests that take the 17..128-sequence path
STB_DECODE_CHAIN=1 run chain_tests 400 python -m pytest tests/test_gpu_decode.py -x -q -k "large_width or large_batch or above_64"
# 4. A/B of the decode step (large-v3 width, 4 decoder layers, 120 windows): ms per step from the slope of two run lengths
run step_base 300 python tools/microbench.py step 120 4
STB_XATTN_V2=1 run step_xattn_v2 300 python tools/microbench.py step 120 4
STB_DECODE_CHAIN=1 run step_chain 300 python tools/microbench.py step 120 4
STB_XATTN_V2=1 STB_DECODE_CHAIN=1 run step_both 300 python tools/microbench.py step 120 4
grep -h "ms_per_step" "$OUT"/step_*.log | cut -c1-160 | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
