#!/bin/bash
# Round-2 GPU call "g": base-2 exponentials in the attention kernels; programmatic dependent launch on / off per kernel class.
O=gpurun_out/r2g
mkdir -p $O
echo "=== decode + depth tests (exp2)" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_depth.py -x -q -m gpu -s > $O/decode_tests.log 2>&1
echo "rc=$? $(tail -1 $O/decode_tests.log)" >> $O/summary.txt
for v in "pdl3:STB_PDL=3" "pdl0:STB_PDL=0" "pdl1:STB_PDL=1" "pdl2:STB_PDL=2" "pdl3b:STB_PDL=3" "pdl0b:STB_PDL=0"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-300)" >> $O/summary.txt
done
echo "=== bench pdl0" >> $O/summary.txt
STB_PDL=0 timeout 900 python bench.py --no-cpu-baseline > $O/bench_pdl0.json 2> $O/bench_pdl0.err
echo "rc=$? $(cut -c1-200 $O/bench_pdl0.json)" >> $O/summary.txt
echo "=== bench pdl3" >> $O/summary.txt
timeout 900 python bench.py > $O/bench_pdl3.json 2> $O/bench_pdl3.err
echo "rc=$? $(cut -c1-200 $O/bench_pdl3.json)" >> $O/summary.txt
cat $O/summary.txt
