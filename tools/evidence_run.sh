set -x
O=gpurun_out/ev
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests exit $?" >> $O/tests.log)
timeout 500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference_arm.json 2> $O/bench_ref.err
timeout 200 python tools/step_breakdown.py > $O/step_breakdown_live.txt 2>&1
timeout 200 python tools/gemm_shapes.py > $O/gemm_shapes_vs_cublas.txt 2>&1
timeout 200 python tools/attn_probe.py > $O/attn_probe.txt 2>&1
timeout 200 python tools/attn_fwd_probe.py --versions v1,v2,v3,v4 --no-check > $O/attn_fwd_probe.txt 2>&1
timeout 200 python tools/adamw_probe.py > $O/adamw_probe.txt 2>&1
# launch list of the bench step (the same command exited 0 above)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 700 --csv --log-file $O/launches_bench_step.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_launches.log 2>&1
# full captures: attention forward v4, attention backward, pair GEMM fc1 dgrad, fc1 gelu
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd4 -s 6 -c 1 -o $O/attn_fwd4 -f python tools/gpu_check.py --one attention > $O/ncu_attn_fwd4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_kernel -s 6 -c 1 -o $O/attn_bwd -f python tools/gpu_check.py --one attention > $O/ncu_attn_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o $O/gemm_pair_fc1_dgrad -f python tools/gpu_check.py --one gemm_dgrad_vitb_fc1 > $O/ncu_gemm1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o $O/gemm_pair_fc1_gelu -f python tools/gpu_check.py --one gemm_fwd_vitb_fc1_gelu > $O/ncu_gemm2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adamw_kernel -s 2 -c 1 -o $O/adamw -f python tools/adamw_probe.py > $O/ncu_adamw.log 2>&1
tail -3 $O/tests.log
python -c "import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['fedavg']['round_ms'])"
cat $O/bench_n1_reference_arm.json | head -c 400
ls -la $O
