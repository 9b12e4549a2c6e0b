mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log
tail -30 gpurun_out/t_pytest.log
