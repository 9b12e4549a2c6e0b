mkdir -p gpurun_out
timeout 400 python tools/fuzz_vs_reference.py --mode constraints --first 30000 --count 4000 --seconds 200 > gpurun_out/fuzz3.json 2> gpurun_out/fuzz3.err; echo "rc=$?"
tail -c 1500 gpurun_out/fuzz3.json; tail -c 600 gpurun_out/fuzz3.err
timeout 400 python tools/fuzz_vs_reference.py --mode pcr --first 40000 --count 4000 --seconds 200 > gpurun_out/fuzz4.json 2> gpurun_out/fuzz4.err; echo "rc=$?"
tail -c 1500 gpurun_out/fuzz4.json; tail -c 600 gpurun_out/fuzz4.err
