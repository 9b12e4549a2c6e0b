mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
for c in edit25_k3 edit20_k2; do
  timeout 300 python tools/run_config.py --config $c --bases 1000000000 >> gpurun_out/f_configs.jsonl 2>> gpurun_out/f_configs.err
done
timeout 900 python tools/cli_timing.py --bases 3000000000 > gpurun_out/f_cli_timing.json 2> gpurun_out/f_cli_timing.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_edit -c 1 -f -o gpurun_out/r2_after4_edit25_scan python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/f_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:index_verify_lv -c 1 -f -o gpurun_out/r2_after4_edit25_verify python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/f_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_edit25.csv python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 1 > gpurun_out/f_ncu3.log 2>&1
tail -3 gpurun_out/f_pytest.log; cat gpurun_out/f_configs.jsonl | cut -c1-200; tail -3 gpurun_out/f_configs.err; cat gpurun_out/f_cli_timing.json | cut -c1-1500; tail -3 gpurun_out/f_cli_timing.err
