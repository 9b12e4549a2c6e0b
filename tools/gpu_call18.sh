mkdir -p gpurun_out
timeout 900 python tools/fuzz_vs_reference.py --first 10000 --count 4000 --seconds 420 > gpurun_out/fuzz1.json 2> gpurun_out/fuzz1.err; echo "rc=$?"
tail -c 1500 gpurun_out/fuzz1.json; tail -c 500 gpurun_out/fuzz1.err
