python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'])
k=d['kernel_ms_per_step']; print({a: k[a] for a in ('usher','usher_eval0','usher_fix','pair')})
print('e2e', d['e2e']['value'], 'plugin', d['e2e_lammps_plugin']['value'])
for s in ('r0','r1_equil300','strong_16m','shear_4m','like_for_like'):
    print(s, round(d[s]['ms_per_step'],4), round(d[s]['value']/1e6,1), d[s]['trials_per_step'], d[s]['kernel_ms_per_step'].get('usher_fix'))
PY
