for p in 102481 204973 394019 788023 1514783 3029419 5823247 11645891 22386109 7240481 14480189; do
  for m in 0 1 2 3; do echo -n "mask=$m "; MARIN_B200_R4MASK=$m python tools/pass_times.py $p 2>&1 | tail -1; done
done > gpurun_out/r2q_lowlat2.log 2>&1
