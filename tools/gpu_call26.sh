mkdir -p gpurun_out
timeout 500 python tools/fuzz_vs_reference.py --first 50000 --count 9000 --seconds 300 > gpurun_out/fuzz7.json 2> gpurun_out/fuzz7.err; echo "rc=$?"
tail -c 900 gpurun_out/fuzz7.json | cut -c1-900; tail -c 300 gpurun_out/fuzz7.err
timeout 300 python tools/fuzz_vs_reference.py --mode constraints --first 60000 --count 9000 --seconds 150 > gpurun_out/fuzz8.json 2> gpurun_out/fuzz8.err; echo "rc=$?"
tail -c 600 gpurun_out/fuzz8.json; tail -c 300 gpurun_out/fuzz8.err
