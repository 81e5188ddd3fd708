#!/bin/bash
# Round-2 GPU call "j": LayerNorm folded into the decode-step linears (option decode_fused_ln) -- parity, A/B, bench.
O=gpurun_out/r2j
mkdir -p $O
echo "=== decode tests (default) incl. the folded-LayerNorm parity test" > $O/summary.txt
timeout 1200 python -m pytest tests/test_gpu_decode.py -q -m gpu -s > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
grep -h "folded LayerNorm" $O/decode_tests.log | head >> $O/summary.txt
echo "=== depth + decode + sampling tests with decode_fused_ln=1" >> $O/summary.txt
STB_DECODE_FUSED_LN=1 timeout 1500 python -m pytest tests/test_gpu_depth.py tests/test_gpu_decode.py tests/test_gpu_sampling.py -q -m gpu -s > $O/fused_tests.log 2>&1
echo "rc=$? $(tail -1 $O/fused_tests.log)" >> $O/summary.txt
grep -h "rel err\|worst" $O/fused_tests.log | head -12 >> $O/summary.txt
for v in "ln0:STB_DECODE_FUSED_LN=0" "ln1:STB_DECODE_FUSED_LN=1" "ln0_b:STB_DECODE_FUSED_LN=0" "ln1_b:STB_DECODE_FUSED_LN=1"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-420)" >> $O/summary.txt
done
echo "=== bench fused" >> $O/summary.txt
STB_DECODE_FUSED_LN=1 timeout 900 python bench.py > $O/bench_fused.json 2> $O/bench_fused.err
echo "rc=$? $(cut -c1-200 $O/bench_fused.json)" >> $O/summary.txt
cat $O/summary.txt
