mkdir -p gpurun_out
md5sum sequence-alignment-tools_b200/lib/libpm_b200.so > gpurun_out/n_md5.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -3 gpurun_out/n_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/n_bench2.json 2> gpurun_out/n_bench2.err; echo "bench rc=$?" >> gpurun_out/n_bench2.err
tail -c 300 gpurun_out/n_bench2.err; grep "^{" gpurun_out/n_bench2.json | cut -c1-300
