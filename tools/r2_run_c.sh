#!/bin/bash
# Round-2 GPU call "c": first hardware run of decode_linear (cluster split-K) and the persistent cross-attention variant.
# Order: the parity tests that exercise the new kernels, the A/B of the decode step, the whole -m gpu suite, the default bench.
mkdir -p gpurun_out
O=gpurun_out/r2c
mkdir -p $O
echo "=== decode tests" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_depth.py -x -q -m gpu -s > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
for v in "base:" "legacy:STB_DECODE_SPLITK_LEGACY=1" "persist:STB_XATTN_PERSIST=1"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-900)" >> $O/summary.txt
done
echo "=== persist parity" >> $O/summary.txt
STB_XATTN_PERSIST=1 timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > $O/persist_tests.log 2>&1
echo "rc=$? $(tail -1 $O/persist_tests.log)" >> $O/summary.txt
echo "=== all gpu tests" >> $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/gputests.log 2>&1
echo "rc=$? $(tail -1 $O/gputests.log)" >> $O/summary.txt
echo "=== bench" >> $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "rc=$? $(cut -c1-300 $O/bench.json)" >> $O/summary.txt
cat $O/summary.txt
