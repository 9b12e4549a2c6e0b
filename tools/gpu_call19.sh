mkdir -p gpurun_out
timeout 300 python tools/dbg/case10573b.py > gpurun_out/dbg2.txt 2>&1; echo rc=$?
cat gpurun_out/dbg2.txt | cut -c1-400
