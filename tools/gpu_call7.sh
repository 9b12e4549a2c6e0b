mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g_smi.txt 2>&1
nvidia-smi topo -m > gpurun_out/g_topo.txt 2>&1
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/g_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/g_bench2.json 2> gpurun_out/g_bench2.err; echo "bench rc=$?" >> gpurun_out/g_bench2.err
tail -3 gpurun_out/g_pytest_multi.log; tail -c 600 gpurun_out/g_bench2.err; cut -c1-600 gpurun_out/g_bench2.json
