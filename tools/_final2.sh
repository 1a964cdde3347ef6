python -m pytest tests -m gpu -x -q > gpurun_out/r2x_gputests.log 2>&1; tail -2 gpurun_out/r2x_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; tail -c 300 gpurun_out/r2x_bench.json
python tools/bench_ladder.py > gpurun_out/r2x_ladder.log 2>&1; tail -12 gpurun_out/r2x_ladder.log | head -11
timeout 400 python tools/mm31_ultralow.py > gpurun_out/r2x_mm31.log 2>&1; tail -2 gpurun_out/r2x_mm31.log
