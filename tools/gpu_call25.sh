mkdir -p gpurun_out
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so > gpurun_out/u_md5.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/u_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/u_smoke.log; tail -2 gpurun_out/u_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
tail -5 gpurun_out/u_pytest.log
timeout 200 python bench.py --config headline --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"
grep "^{" gpurun_out/u_bench.json | cut -c1-260
