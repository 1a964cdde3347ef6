python tools/quick_time.py 86243 102481 172001 204973 350003 394019 700001 788023 1362763 1514783 2500009 3029419 4999999 5823247 > gpurun_out/r2r_quick.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2r_gputests.log 2>&1
tail -3 gpurun_out/r2r_gputests.log; grep "us/iter" gpurun_out/r2r_quick.log
