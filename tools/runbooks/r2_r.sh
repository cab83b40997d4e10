# Round 2, closing call (one B200) at HEAD: the suite, smoke(), the default bench line, BASELINE configs[4] as one JSON
# line, and -- each after its plain run has exited 0 -- the ncu launch list of the bench command and the `--set full`
# capture of the three passes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2r_tests.log 2>&1; tail -4 gpurun_out/r2r_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; tail -1 gpurun_out/r2r_smoke.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench.json') if l.startswith('{')][-1])
r=d['roofline']; print('value',d['value'],'ms',d['ms_per_step'],'frac',r['frac'],r['matmult']['frac'],{k:round(v['ms'],4) for k,v in r['passes'].items()},'traffic',r['traffic'],r.get('traffic_source'))
print('cg',d['cg']['its'],d['cg']['time_s'],d['cg']['gpu_launches'],'e2e',d['e2e']['value'],d['e2e'].get('batch',{}).get('value'),d['e2e'].get('cg'))
print('parity',d['parity']['ok'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline'].get('cg_iteration_parity'), d['clocks'])
PY
timeout 900 python bench.py --workload config5 > gpurun_out/r2r_config5.json 2> gpurun_out/r2r_config5.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2r_config5.json') if l.startswith('{')][-1]); print(d['frac_range'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --quick --no-parity --cg-maxit 3 > gpurun_out/r2r_launches.log 2>&1
tail -2 gpurun_out/r2r_launches.log | cut -c1-200
timeout 120 python tools/prof_lapl.py --n 512 --reps 2 > gpurun_out/r2r_prof_lapl.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 3 -f -o gpurun_out/r2r_full_512 \
  python tools/prof_lapl.py --n 512 --reps 2 > gpurun_out/r2r_full_512.log 2>&1
tail -2 gpurun_out/r2r_prof_lapl.log
