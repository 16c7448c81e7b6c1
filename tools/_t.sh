python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 100 --warmup 5 --equil 2000 --no-cpu --e2e-steps 0 --lammps-steps 0 --secondary-steps 30"
for cfg in "OBMD_B200_X=1" "OBMD_B200_USH_PLANES=1"; do
echo "== $cfg"; env $cfg $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['kernel_ms_per_step']
        print(d['ms_per_step'], {a: k[a] for a in ('usher','usher_eval0','usher_fix','pair','sort','host_step')}, d['usher']['trials_per_step'])
        for s in ('r1_equil300','strong_16m','shear_4m','like_for_like'):
            print(s, d[s]['ms_per_step'], d[s]['trials_per_step'], d[s]['kernel_ms_per_step'].get('usher_eval0'))
"
done
