#!/bin/bash
# Round-2 GPU call "h": fp16 self-attention K/V cache; tensor-core cross-attention (option xattn_tc) vs the scalar kernel.
O=gpurun_out/r2h
mkdir -p $O
echo "=== decode + depth + sampling tests" > $O/summary.txt
timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_depth.py tests/test_gpu_sampling.py -x -q -m gpu -s > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
grep -h "rel\|tensor-core\|worst" $O/decode_tests.log | head -30 >> $O/summary.txt
echo "=== depth tests with xattn_tc" >> $O/summary.txt
STB_XATTN_TC=1 timeout 900 python -m pytest tests/test_gpu_depth.py tests/test_gpu_decode.py -x -q -m gpu > $O/tc_tests.log 2>&1
echo "rc=$? $(tail -1 $O/tc_tests.log)" >> $O/summary.txt
for v in "scalar:STB_XATTN_TC=0" "tc:STB_XATTN_TC=1" "scalar_b:STB_XATTN_TC=0" "tc_b:STB_XATTN_TC=1"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-330)" >> $O/summary.txt
done
echo "=== bench tc" >> $O/summary.txt
STB_XATTN_TC=1 timeout 900 python bench.py --no-cpu-baseline > $O/bench_tc.json 2> $O/bench_tc.err
echo "rc=$? $(cut -c1-200 $O/bench_tc.json)" >> $O/summary.txt
echo "=== bench scalar" >> $O/summary.txt
timeout 900 python bench.py > $O/bench_scalar.json 2> $O/bench_scalar.err
echo "rc=$? $(cut -c1-200 $O/bench_scalar.json)" >> $O/summary.txt
cat $O/summary.txt
