#!/bin/bash
# SUPERSEDED by tools/r2_final3.sh: this first attempt kept three `ncu --set full` reports (71 MB) under gpurun_out/ and
# gpurun returns at most 64 MiB, so nothing came back but the stdout tail (profiles/r2_final_call1_stdout.txt).
# Round-2 final GPU call: the whole -m gpu suite, the headline bench + the reference arm, the other BASELINE configurations
# (base single clip, small align x64, large-v3 refine), the ncu launch list of a (shortened) step and full captures of the
# heaviest kernels.
O=gpurun_out/r2final
mkdir -p $O
echo "=== all gpu tests" > $O/summary.txt
timeout 1800 python -m pytest tests -q -m gpu > $O/gputests.log 2>&1
echo "rc=$? $(tail -1 $O/gputests.log)" >> $O/summary.txt
echo "=== smoke" >> $O/summary.txt
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
echo "rc=$? $(grep smoke: $O/smoke.log | tr '\n' ' ' | cut -c1-300)" >> $O/summary.txt
echo "=== bench (headline: large-v3 transcribe, 120 windows)" >> $O/summary.txt
timeout 1200 python bench.py > $O/bench_largev3_w120.json 2> $O/bench_largev3_w120.err
echo "rc=$? $(cut -c1-200 $O/bench_largev3_w120.json)" >> $O/summary.txt
echo "=== bench --impl reference" >> $O/summary.txt
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "rc=$? $(cut -c1-400 $O/bench_reference.json)" >> $O/summary.txt
echo "=== config 2: base, one 30 s clip" >> $O/summary.txt
timeout 600 python bench.py --model base --windows 1 --steps 10 --warmup 3 > $O/bench_base_w1.json 2> $O/bench_base_w1.err
echo "rc=$? $(cut -c1-200 $O/bench_base_w1.json)" >> $O/summary.txt
echo "=== config 3: small, align, 64 windows" >> $O/summary.txt
timeout 600 python bench.py --model small --workload align --windows 64 --steps 10 --warmup 3 > $O/bench_small_align_w64.json 2> $O/bench_small_align_w64.err
echo "rc=$? $(cut -c1-200 $O/bench_small_align_w64.json)" >> $O/summary.txt
echo "=== config 5: large-v3 refine, 3 groups" >> $O/summary.txt
timeout 900 python bench.py --workload refine --steps 6 --warmup 3 > $O/bench_largev3_refine.json 2> $O/bench_largev3_refine.err
echo "rc=$? $(cut -c1-200 $O/bench_largev3_refine.json)" >> $O/summary.txt
echo "=== ncu launch list (large-v3 transcribe, 120 windows, 16 decode steps)" >> $O/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_largev3_w120_t16.csv \
  python bench.py --ncu --tokens 16 > $O/ncu_launches.log 2>&1
echo "rc=$? $(wc -l < $O/launches_largev3_w120_t16.csv) lines" >> $O/summary.txt
echo "=== ncu full: encoder GEMM / attention / forced-pass kernels (large-v3, 8 windows)" >> $O/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'gemm_tc_kernel|attention_tc_kernel|softmax_kernel|layernorm_kernel' --launch-skip 20 -c 14 -o $O/ncu_encoder -f \
  python bench.py --ncu --tokens 16 --windows 8 > $O/ncu_encoder.log 2>&1
echo "rc=$? $(ls -la $O/ncu_encoder.ncu-rep 2>&1 | cut -c1-100)" >> $O/summary.txt
echo "=== ncu full: DTW / QK post-processing (small, align, 64 windows)" >> $O/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'dtw|qk_|token_prob|logmel' -c 12 -o $O/ncu_align -f \
  python bench.py --ncu --model small --workload align --windows 64 > $O/ncu_align.log 2>&1
echo "rc=$? $(ls -la $O/ncu_align.ncu-rep 2>&1 | cut -c1-100)" >> $O/summary.txt
echo "=== ncu full: decode-step kernels" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_linear_kernel|decode_cross_attn|decode_self_attn_kernel|sample_greedy' \
  --launch-skip 400 -c 16 -o $O/ncu_step -f python tools/microbench.py step 120 2 > $O/ncu_step.log 2>&1
echo "rc=$? $(ls -la $O/ncu_step.ncu-rep 2>&1 | cut -c1-100)" >> $O/summary.txt
cat $O/summary.txt
