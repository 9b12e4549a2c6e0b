mkdir -p gpurun_out
timeout 170 python tools/fuzz_vs_reference.py --long --first 310000 --count 3000 --seconds 100 > gpurun_out/fuzz9.json 2> gpurun_out/fuzz9.err; echo "rc=$?"
tail -c 700 gpurun_out/fuzz9.json | cut -c1-700; tail -c 300 gpurun_out/fuzz9.err
