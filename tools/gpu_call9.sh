mkdir -p gpurun_out
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so > gpurun_out/i_md5.txt; strings sequence-alignment-tools_b200/lib/libpm_b200.so | grep "k out of range" >> gpurun_out/i_md5.txt; ls -la --time-style=full-iso sequence-alignment-tools_b200/lib/ sequence-alignment-tools_b200/csrc/pm_b200.cu >> gpurun_out/i_md5.txt
timeout 1200 python -m pytest tests/test_k45.py tests/test_gpu_parity.py tests/test_large_golden.py -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so >> gpurun_out/i_md5.txt
for c in edit25_k3 edit20_k2 p100_k1; do
  timeout 300 python tools/run_config.py --config $c --bases 1000000000 >> gpurun_out/i_configs.jsonl 2>> gpurun_out/i_configs.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_edit -c 1 -f -o gpurun_out/r2_after6_edit25_scan python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 1 --warmup 0 > gpurun_out/i_ncu1.log 2>&1
cat gpurun_out/i_md5.txt; tail -3 gpurun_out/i_pytest.log; cat gpurun_out/i_configs.jsonl | cut -c1-160
