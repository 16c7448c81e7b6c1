python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "usher or pre_exchange or b3 or b8 or many_insertions" > gpurun_out/$1_tests.log 2>&1; tail -3 gpurun_out/$1_tests.log
python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu --e2e-steps 0 > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err
python - <<PY
import json
for l in open("gpurun_out/$1_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]
        print("ms/step", round(d["ms_per_step"],4), {x:k[x] for x in ("usher","usher_eval0","usher_fix","pair","pre_exchange","sort","bin","nve","post_force")}, d["usher"]["trials_per_step"], d["usher"]["rounds_per_step"])
PY
