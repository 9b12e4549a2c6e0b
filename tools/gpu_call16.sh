mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fasta_device.py tests/test_gpu_cli.py tests/test_sqz.py tests/test_adapter.py -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
tail -5 gpurun_out/o_pytest.log
PM_TIMING=1 python tools/cli_timing.py > gpurun_out/o_cli_timing.json 2> gpurun_out/o_cli_timing.err; echo "cli rc=$?"
tail -c 600 gpurun_out/o_cli_timing.json
