#!/bin/bash
# Round-2 GPU call "e": batch stepped as 2 / 3 concurrent chains on separate streams (DualStepEngine) vs one chain.
O=gpurun_out/r2e
mkdir -p $O
echo "=== decode tests (dual default)" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_sampling.py tests/test_gpu_depth.py -x -q -m gpu > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
for v in "dual0:STB_DECODE_DUAL=0" "dual2:STB_DECODE_DUAL=2" "dual3:STB_DECODE_DUAL=3" "dual4:STB_DECODE_DUAL=4" "dual2_noprio:STB_DECODE_DUAL=2 STB_DECODE_LIN_PRIORITY=0" "dual2_legacy:STB_DECODE_DUAL=2 STB_DECODE_SPLITK_LEGACY=1"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-200)" >> $O/summary.txt
done
echo "=== bench dual2" >> $O/summary.txt
timeout 900 python bench.py > $O/bench_dual2.json 2> $O/bench_dual2.err
echo "rc=$? $(cut -c1-200 $O/bench_dual2.json)" >> $O/summary.txt
echo "=== bench dual0" >> $O/summary.txt
STB_DECODE_DUAL=0 timeout 900 python bench.py --no-cpu-baseline > $O/bench_dual0.json 2> $O/bench_dual0.err
echo "rc=$? $(cut -c1-200 $O/bench_dual0.json)" >> $O/summary.txt
cat $O/summary.txt
