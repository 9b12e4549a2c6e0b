mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -q -k "uppercase" > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -40 gpurun_out/p_pytest.log
