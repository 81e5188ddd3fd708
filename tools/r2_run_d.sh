#!/bin/bash
# Round-2 GPU call "d": decode_linear with the vectorised DSMEM reduction, sampling / ragged-prompt kernels, timing variants.
O=gpurun_out/r2d
mkdir -p $O
echo "=== decode + sampling tests" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_sampling.py -x -q -m gpu -s > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
for v in "base:" "legacy:STB_DECODE_SPLITK_LEGACY=1"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-900)" >> $O/summary.txt
done
echo "=== all gpu tests" >> $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/gputests.log 2>&1
echo "rc=$? $(tail -1 $O/gputests.log)" >> $O/summary.txt
echo "=== ncu full: decode-step kernels" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_linear_kernel|decode_cross_attn_kernel|decode_self_attn_kernel' \
  --launch-skip 400 -c 16 -o $O/ncu_step -f python tools/microbench.py step 120 2 > $O/ncu_step.log 2>&1
echo "rc=$? $(ls -la $O/ncu_step.ncu-rep 2>&1 | cut -c1-120)" >> $O/summary.txt
echo "=== bench" >> $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "rc=$? $(cut -c1-300 $O/bench.json)" >> $O/summary.txt
cat $O/summary.txt
