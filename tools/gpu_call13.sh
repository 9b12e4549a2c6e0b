mkdir -p gpurun_out
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so > gpurun_out/l_md5.txt
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -3 gpurun_out/l_pytest.log
