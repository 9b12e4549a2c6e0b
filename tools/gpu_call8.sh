mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?" >> gpurun_out/h_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err
for c in edit25_k3 edit20_k2 headline; do
  timeout 300 python tools/run_config.py --config $c --bases 1000000000 >> gpurun_out/h_configs.jsonl 2>> gpurun_out/h_configs.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_sub -c 1 -f -o gpurun_out/r2_headline_scan python tools/run_config.py --config headline --bases 3000000000 --steps 1 --warmup 0 > gpurun_out/h_ncu0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_edit -c 1 -f -o gpurun_out/r2_after5_edit25_scan python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/h_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:index_verify_lv -c 1 -f -o gpurun_out/r2_after5_edit25_verify python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/h_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_edit25.csv python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 1 > gpurun_out/h_ncu3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config headline > gpurun_out/h_ncu4.log 2>&1
PM_READER_THREADS=16 timeout 900 python tools/cli_timing.py --bases 3000000000 > gpurun_out/h_cli_timing16.json 2> gpurun_out/h_cli_timing16.err
tail -3 gpurun_out/h_pytest.log; tail -c 300 gpurun_out/h_bench.err; cat gpurun_out/h_configs.jsonl | cut -c1-160
