mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/j_topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/j_bench8.json 2> gpurun_out/j_bench8.err; echo "bench rc=$?" >> gpurun_out/j_bench8.err
tail -c 300 gpurun_out/j_bench8.err; grep "^{" gpurun_out/j_bench8.json | cut -c1-300
