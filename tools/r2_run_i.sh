#!/bin/bash
# Round-2 GPU call "i": persistent double-buffered tensor-core cross-attention (xattn_tc = 2) vs the one-item kernel (1).
O=gpurun_out/r2i
mkdir -p $O
echo "=== decode tests (default xattn_tc=1)" > $O/summary.txt
timeout 1200 python -m pytest tests/test_gpu_decode.py -q -m gpu -s > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
grep -h "tensor-core" $O/decode_tests.log | head >> $O/summary.txt
echo "=== depth + decode tests with xattn_tc=2" >> $O/summary.txt
STB_XATTN_TC=2 timeout 1200 python -m pytest tests/test_gpu_depth.py tests/test_gpu_decode.py -q -m gpu > $O/tc2_tests.log 2>&1
echo "rc=$? $(tail -1 $O/tc2_tests.log)" >> $O/summary.txt
for v in "tc1:STB_XATTN_TC=1" "tc2:STB_XATTN_TC=2" "tc1_b:STB_XATTN_TC=1" "tc2_b:STB_XATTN_TC=2"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-330)" >> $O/summary.txt
done
echo "=== bench tc2" >> $O/summary.txt
STB_XATTN_TC=2 timeout 900 python bench.py --no-cpu-baseline > $O/bench_tc2.json 2> $O/bench_tc2.err
echo "rc=$? $(cut -c1-200 $O/bench_tc2.json)" >> $O/summary.txt
cat $O/summary.txt
