#!/bin/bash
# Round-2 closing GPU call: (1) the whole -m gpu suite on the final library (no restrict-qualified kernel parameters);
# (2) LayerNorm folded into the decode-step linears: parity with the option on, A/B of the step; (3) the artefacts (bench lines
# of every BASELINE configuration, ncu launch list, ncu --set full captures exported to CSV on the box) with the faster,
# parity-clean setting of (2) exported as STB_DECODE_FUSED_LN.
O=gpurun_out/r2final3
mkdir -p $O
echo "=== all gpu tests" > $O/summary.txt
timeout 1800 python -m pytest tests -q -m gpu -s > $O/gputests.log 2>&1
rc_all=$?
echo "rc=$rc_all $(tail -1 $O/gputests.log)" >> $O/summary.txt
grep -h "folded LayerNorm\|FAILED" $O/gputests.log | head -20 >> $O/summary.txt
echo "=== depth + decode + sampling tests with decode_fused_ln=1" >> $O/summary.txt
STB_DECODE_FUSED_LN=1 timeout 1500 python -m pytest tests/test_gpu_depth.py tests/test_gpu_decode.py tests/test_gpu_sampling.py -q -m gpu -s > $O/fused_tests.log 2>&1
rc_fused=$?
echo "rc=$rc_fused $(tail -1 $O/fused_tests.log)" >> $O/summary.txt
grep -h "parity mode\|forced 48\|FAILED" $O/fused_tests.log | head -12 >> $O/summary.txt
for v in "ln0:STB_DECODE_FUSED_LN=0" "ln1:STB_DECODE_FUSED_LN=1"; do
  name=${v%%:*}; envs=${v#*:}
  echo "=== step_$name" >> $O/summary.txt
  env $envs timeout 400 python tools/microbench.py step 120 4 > $O/step_$name.log 2>&1
  echo "rc=$? $(tail -1 $O/step_$name.log | cut -c1-420)" >> $O/summary.txt
done
FUSED=0
if [ $rc_fused -eq 0 ] && [ $rc_all -eq 0 ]; then
  FUSED=$(python - <<PY
import json
def ms(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])["step"]["ms_per_step"]
    except Exception:
        return 1e9
a, b = ms("$O/step_ln0.log"), ms("$O/step_ln1.log")
print(1 if 0 < b < a else 0)
PY
)
fi
export STB_DECODE_FUSED_LN=$FUSED
echo "=== chosen: STB_DECODE_FUSED_LN=$FUSED" >> $O/summary.txt
echo "=== bench (headline: large-v3 transcribe, 120 windows)" >> $O/summary.txt
timeout 1200 python bench.py > $O/bench_largev3_w120.json 2> $O/bench_largev3_w120.err
echo "rc=$? $(cut -c1-160 $O/bench_largev3_w120.json)" >> $O/summary.txt
echo "=== config 2: base, one 30 s clip" >> $O/summary.txt
timeout 600 python bench.py --model base --windows 1 --steps 10 --warmup 3 > $O/bench_base_w1.json 2> $O/bench_base_w1.err
echo "rc=$? $(cut -c1-160 $O/bench_base_w1.json)" >> $O/summary.txt
echo "=== config 3: small, align, 64 windows" >> $O/summary.txt
timeout 600 python bench.py --model small --workload align --windows 64 --steps 10 --warmup 3 > $O/bench_small_align_w64.json 2> $O/bench_small_align_w64.err
echo "rc=$? $(cut -c1-160 $O/bench_small_align_w64.json)" >> $O/summary.txt
echo "=== config 5: large-v3 refine, 3 groups" >> $O/summary.txt
timeout 900 python bench.py --workload refine --steps 6 --warmup 3 > $O/bench_largev3_refine.json 2> $O/bench_largev3_refine.err
echo "rc=$? $(cut -c1-160 $O/bench_largev3_refine.json)" >> $O/summary.txt
echo "=== bench --impl reference (1 step)" >> $O/summary.txt
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
echo "rc=$? $(cut -c1-160 $O/bench_reference.json)" >> $O/summary.txt
echo "=== ncu launch list (large-v3 transcribe, 120 windows, 16 decode steps)" >> $O/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_largev3_w120_t16.csv \
  python bench.py --ncu --tokens 16 > $O/ncu_launches.log 2>&1
echo "rc=$? $(wc -l < $O/launches_largev3_w120_t16.csv) lines" >> $O/summary.txt
cap() {   # name, kernel regex, count, skip, command...
  local name=$1 regex=$2 count=$3 skip=$4; shift 4
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"$regex" --launch-skip $skip -c $count \
    -o /tmp/$name -f "$@" > $O/$name.log 2>&1
  local rc=$?
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>> $O/$name.log
  echo "rc=$rc $(wc -l < $O/${name}_raw.csv) csv lines" >> $O/summary.txt
  rm -f /tmp/$name.ncu-rep
}
echo "=== ncu full: encoder GEMM / attention / forced-pass kernels (large-v3, 8 windows)" >> $O/summary.txt
cap ncu_encoder 'gemm_tc_kernel|attention_tc_kernel|softmax_kernel|layernorm_kernel' 14 20 python bench.py --ncu --tokens 16 --windows 8
echo "=== ncu full: DTW / QK post-processing (small, align, 64 windows)" >> $O/summary.txt
cap ncu_align 'dtw|qk_|token_prob|logmel' 12 0 python bench.py --ncu --model small --workload align --windows 64
echo "=== ncu full: decode-step kernels (large-v3 width, 120 windows)" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none -k regex:'decode_linear_kernel|decode_cross_attn|decode_self_attn_kernel|sample_greedy' \
  --launch-skip 400 -c 16 -o /tmp/ncu_step -f python tools/microbench.py step 120 2 > $O/ncu_step.log 2>&1
echo "rc=$?" >> $O/summary.txt
ncu -i /tmp/ncu_step.ncu-rep --page raw --csv > $O/ncu_step_raw.csv 2>> $O/ncu_step.log
rm -f /tmp/ncu_step.ncu-rep
du -sh $O >> $O/summary.txt
cat $O/summary.txt
