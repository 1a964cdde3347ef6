for p in 43003 86243 172001 350003 700001 1362763; do python tools/pass_times.py $p; done > gpurun_out/r2q_small_pass.log 2>&1
cat gpurun_out/r2q_small_pass.log
