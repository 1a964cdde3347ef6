python -m pytest tests -m gpu -x -q > gpurun_out/r2v_gputests.log 2>&1; tail -2 gpurun_out/r2v_gputests.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -c 400 gpurun_out/r2v_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_ref.json 2>> gpurun_out/r2v_bench.err
python bench.py --workload pm1 > gpurun_out/r2v_pm1.json 2>> gpurun_out/r2v_bench.err
python bench.py --workload ecm > gpurun_out/r2v_ecm.json 2>> gpurun_out/r2v_bench.err
python tools/bench_ladder.py > gpurun_out/r2v_ladder.log 2>&1; tail -28 gpurun_out/r2v_ladder.log | head -27
python tools/op_times.py > gpurun_out/r2v_op_times.log 2>&1
export QT_COUNT=12
python tools/quick_time.py 136279841 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 24 --csv --log-file gpurun_out/r2v_launches.csv python tools/quick_time.py 136279841 > gpurun_out/ncu1.log 2>&1
python tools/quick_time.py 136279841 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_mid_fast|k_inv_fast' -s 100 -c 2 -o gpurun_out/r2v_prof python tools/quick_time.py 136279841 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/r2v*
