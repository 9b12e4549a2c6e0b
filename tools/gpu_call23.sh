mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_adapter.py -m gpu -q -k "pcr or adapter or halves or constraints" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -25 gpurun_out/s_pytest.log
timeout 400 python tools/fuzz_vs_reference.py --mode pcr --first 41058 --count 6000 --seconds 200 > gpurun_out/fuzz6.json 2> gpurun_out/fuzz6.err; echo "rc=$?"
tail -c 1500 gpurun_out/fuzz6.json; tail -c 600 gpurun_out/fuzz6.err
