# Round 2, last call (one B200): the reduction tail fused into the rotated z kernel too (5 launches per CG iteration at
# 512^3 on one GPU instead of 7) -- suite and the default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2t_tests.log 2>&1; tail -3 gpurun_out/r2t_tests.log | cut -c1-200
timeout 600 python bench.py --quick > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench.json') if l.startswith('{')][-1])
r=d['roofline']; print('value',d['value'],'ms',d['ms_per_step'],'frac',r['frac'],r['matmult']['frac'],{k:round(v['ms'],4) for k,v in r['passes'].items()})
print('cg',d['cg']['its'],d['cg']['time_s'],d['cg']['gpu_launches'],d['cg']['true_residual_rel'],'e2e',d['e2e']['value'], d['clocks'])
PY
