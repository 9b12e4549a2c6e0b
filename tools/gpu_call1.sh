mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
for c in edit25_k3 oligo100k_k1 iupac_k2 headline; do timeout 300 python tools/run_config.py --config $c --bases 1000000000 >> gpurun_out/a_configs.jsonl 2>> gpurun_out/a_configs.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_kernel -c 1 -f -o gpurun_out/r2_before_edit25_scan python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/a_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:index_verify -c 1 -f -o gpurun_out/r2_before_edit25_verify python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/a_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_sub -c 1 -f -o gpurun_out/r2_before_oligo100k_k1 python tools/run_config.py --config oligo100k_k1 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/a_ncu3.log 2>&1
tail -3 gpurun_out/a_pytest.log; tail -c 600 gpurun_out/a_bench.err; cat gpurun_out/a_configs.jsonl | cut -c1-300
