python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'])
print(d['rooflines'])
print('e2e', d['e2e']['value'], 'plugin', d['e2e_lammps_plugin']['value'], 'cpu', d['cpu_baseline']['value'])
for k in ('r0','r1_equil300','strong_16m','shear_4m','like_for_like'):
    print(k, round(d[k]['ms_per_step'],4), round(d[k]['value']/1e6,1), d[k]['trials_per_step'])
PY
