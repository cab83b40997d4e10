# Round 2, call 9 (one B200): the suite at HEAD, determinism of the default paths on the bricks that failed, the default
# bench line, BASELINE configs[4] as one JSON line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2i_tests.log 2>&1; tail -6 gpurun_out/r2i_tests.log
REPS=30 timeout 600 python tools/determinism_check.py 48,640,1088 32,640,1088 512,512,512 > gpurun_out/r2i_determinism.log 2>&1; cat gpurun_out/r2i_determinism.log
timeout 600 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench.json') if l.startswith('{')][-1])
r=d['roofline']; print('value',d['value'],'ms',d['ms_per_step'],'frac',r['frac'],r['matmult']['frac'],{k:round(v['ms'],4) for k,v in r['passes'].items()},'traffic',r['traffic'],r.get('traffic_source'))
print('cg',d['cg']['its'],d['cg']['time_s'],d['cg']['gpu_launches'],'e2e',d['e2e']['value'],d['e2e'].get('batch',{}).get('value'),d['e2e'].get('cg'))
print('parity',d['parity']['ok'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline'].get('cg_iteration_parity'))
PY
timeout 900 python bench.py --workload config5 > gpurun_out/r2i_config5.json 2> gpurun_out/r2i_config5.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2i_config5.json') if l.startswith('{')][-1]); print(d['frac_range'])"
