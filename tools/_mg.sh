for n in 8 4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 --no-secondary --no-cpu --e2e-steps 0 --lammps-steps 0 > gpurun_out/r2_bench_n${n}_default.json 2> gpurun_out/r2_bench_n${n}_default.err; echo "bench n$n rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${n}_default.json').read().strip().splitlines()[-1])
    print($n, d['ms_per_step'], d['value'], d['usher']['trials_per_step'], d['kernel_ms_per_step'], d.get('slab_check'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r2_bench_n${n}_default.err').read()[-1500:])
PY
done
