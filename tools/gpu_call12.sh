mkdir -p gpurun_out
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so > gpurun_out/k_md5.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/k_smoke.log
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?" >> gpurun_out/k_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/k_bench_ref.json 2> gpurun_out/k_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config headline > gpurun_out/k_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_sub -c 1 -f -o gpurun_out/r2_headline_scan python tools/run_config.py --config headline --bases 3000000000 --steps 1 --warmup 0 > gpurun_out/k_ncu0.log 2>&1
tail -2 gpurun_out/k_smoke.log; tail -3 gpurun_out/k_pytest.log; tail -c 200 gpurun_out/k_bench.err
