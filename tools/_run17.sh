# end-of-round validation: the whole GPU suite, smoke, the headline bench line, the ECM line
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_gputests.log 2>&1; tail -3 gpurun_out/r2z_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1]); print('prp', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d['clocks'])"
python bench.py --workload ecm --steps 3 > gpurun_out/r2z_ecm.json 2> gpurun_out/r2z_ecm.err; python -c "
import json; d=json.loads(open('gpurun_out/r2z_ecm.json').read().strip().splitlines()[-1]); print('ecm', round(d['value'],2), d['unit'], 'us/point-op', round(d['us_per_point_op_launch'],1), 'seq', d['sequential_one_engine_per_curve']['curves_per_s'], 'big', d['saturating_batch'])"
