mkdir -p gpurun_out
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so > gpurun_out/m_md5.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tandem or seed_list" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -30 gpurun_out/m_pytest.log
