mkdir -p gpurun_out
timeout 300 python tools/dbg/case10573.py > gpurun_out/dbg3.txt 2>&1; echo rc=$?
grep -c "index cands \[\]" gpurun_out/dbg3.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "flagged or slow_path or seam or across_tiles" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -15 gpurun_out/q_pytest.log
timeout 900 python tools/fuzz_vs_reference.py --first 20000 --count 4000 --seconds 240 > gpurun_out/fuzz2.json 2> gpurun_out/fuzz2.err; echo "rc=$?"
tail -c 1200 gpurun_out/fuzz2.json; tail -c 300 gpurun_out/fuzz2.err
