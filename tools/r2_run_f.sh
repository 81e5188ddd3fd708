#!/bin/bash
# Round-2 GPU call "f": uniform shared-memory carveout across the decode-step kernels; 16 key splits in the cross-attention.
O=gpurun_out/r2f
mkdir -p $O
echo "=== decode tests" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
echo "=== decode tests, 16 splits" >> $O/summary.txt
STB_XATTN_SPLITS=16 timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > $O/decode_tests_s16.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests_s16.log)" >> $O/summary.txt
for v in "base:" "nocarve:STB_UNIFORM_CARVEOUT=0" "s16:STB_XATTN_SPLITS=16" "nopdl:STB_PDL=0"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-700)" >> $O/summary.txt
done
echo "=== bench" >> $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "rc=$? $(cut -c1-200 $O/bench.json)" >> $O/summary.txt
echo "=== bench s16" >> $O/summary.txt
STB_XATTN_SPLITS=16 timeout 900 python bench.py --no-cpu-baseline > $O/bench_s16.json 2> $O/bench_s16.err
echo "rc=$? $(cut -c1-200 $O/bench_s16.json)" >> $O/summary.txt
cat $O/summary.txt
