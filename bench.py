#!/usr/bin/env python
"""bench.py -- OBMD-DPD atom-steps/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path

A "step" is one MD timestep of the whole hot path (FixNVE halves, cell build, fix obmd
pre_exchange incl. batched USHER, PairDPD, fix obmd post_force) over every atom.  Workload at
N=1: BASELINE config 2, "OBMD-DPD fluid at equilibrium, 1M particles" = the shipped
configuration tiled 4x5x4 (1 010 960 atoms), steady-exchange regime R1 (alpha 1.0,
nbuf = int(0.15 N)), see SURVEY 8(d).  value = sum over steps of the atom count / device time
(CUDA events on the library's stream, max over ranks).  e2e = the same metric through the C ABI
with HOST-owned atom arrays: every step uploads x,v from pinned memory (obmd_upload_xv), runs one
device step (obmd_run) and reads x,v,f,tag,type,image back into pinned memory (obmd_download).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner / INFO lines go to stderr: stdout is one JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "OBMD-DPD atom-steps/sec"
UNIT = "atom-steps/s"
PAIR_BYTES_PER_ATOM = 80.0   # x,v,tag,type in + f out
USHER = (1, 31.03, 1.0, 0.02, 1e4, 1.5, 1.0, 40)


def shipped():
    d = np.load(os.path.join(ROOT, "tests", "golden", "dpd_8map_obmd.npz"))
    return d["x"].copy(), d["box"].copy()


def tiled(reps):
    x, box = shipped()
    L = box[:, 1] - box[:, 0]
    parts = [x + np.array([i, j, k]) * L for i in range(reps[0]) for j in range(reps[1])
             for k in range(reps[2])]
    X = np.concatenate(parts)
    B = np.array([[0.0, L[0] * reps[0]], [0.0, L[1] * reps[1]], [0.0, L[2] * reps[2]]])
    return X, B


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons of this rank's GPU during the timed region
    (B200_PROFILING.md recipe).  Sampled in-process through NVML (nvidia-ml-py): spawning
    `nvidia-smi` from every rank inside the timed region stalls the driver for milliseconds at a
    time, which a 1 ms step with several host round trips cannot hide.  Falls back to one
    nvidia-smi process per sample when NVML is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + uuid)
            self.nvml = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            # the first queries of a process are slow (tens of ms) and serialise with kernel launches in
            # the driver: pay for them here, before the timed region starts
            for _ in range(3):
                self.sample_nvml()
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bit = lambda name, alt: getattr(n, name, getattr(n, alt, 0))
        flags = [bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 bit("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 bit("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")]
        return [str(sm), str(self.max_sm), str(pw)] + ["Active" if (f and (r & f)) else "Not Active" for f in flags]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout.strip()
                    if out:
                        self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05 if self.nvml else 0.25)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows),
                "reasons": reasons, "source": "nvml" if self.nvml else "nvidia-smi"}


def build_sim(x, box, device, regime="R1", temperature=1.0, rank=0, world=1, buffer_size=None, nbuf_atoms=None):
    """Context for this rank.  world > 1: x-slab decomposition (obmd_comm_init), every rank uploads
    only the atoms inside its slab; the NCCL unique id travels through torch.distributed."""
    import common_cases as cc
    import obmd_b200 as ob
    n = len(x)
    nb = nbuf_atoms if nbuf_atoms is not None else n
    if regime == "R1":
        kw = dict(alpha=1.0, nbuf=float(int(0.15 * nb)))
    else:
        kw = dict(alpha=0.7, nbuf=float(int(0.7 * 0.15 * nb)))
    if buffer_size is not None:
        kw["buffer_size"] = buffer_size
    sim = cc.device_sim(box, device=device, temperature=temperature, **kw)
    tag = np.arange(1, n + 1, dtype=np.int32)
    if world > 1:
        import torch.distributed as dist
        uid = [ob.ObmdB200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.comm_init(rank, world, uid[0])
        lo, hi = sim.comm_slab()
        mine = (x[:, 0] >= lo) & ((x[:, 0] < hi) | ((rank == world - 1) & (x[:, 0] <= box[0, 1])))
        x, tag = x[mine], tag[mine]
    m = len(x)
    sim.upload(x, np.zeros_like(x), np.ones(m, np.int32), tag, capacity=int(1.15 * m) + 65536)
    sim.setup(0)
    return sim, kw


def ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not os.environ.get("OBMD_BENCH_NO_PIN"):
        # one disjoint set of cores per rank (as `numactl`/`taskset` would give each MPI rank): the step
        # has several host round trips, and rank threads migrating onto each other's cores showed up as
        # sporadic 30-100 % slow-downs of whole runs on the virtualised GPU hosts
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local * per:(local + 1) * per] or cores
            os.sched_setaffinity(0, mine)
        except (AttributeError, OSError):
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # weak scaling: the box grows along the open axis with the number of GPUs (one 4x5x4 tiling =
    # 1 010 960 atoms per GPU), the buffers keep the physical length and target population of N=1
    scaling = "weak"
    if args.workload == "1m":
        reps1 = tuple(args.reps)
        reps = (reps1[0] * world, reps1[1], reps1[2])
        B1 = 0.15 * 33.59462486002239 * reps1[0]
    else:
        # BASELINE configs 3-5: fixed total size, slab-decomposed over the GPUs (strong scaling).
        # The buffer is 0.15 Lx unless that exceeds what one edge slab can hold.
        reps = {"4m": (5, 8, 8), "16m": (20, 8, 8)}[args.workload]
        scaling = "strong"
        Lx = 33.59462486002239 * reps[0]
        B1 = min(0.15 * Lx, Lx / world - 2.5)
    x, box = tiled(reps)
    natoms_total = len(x)
    n0 = natoms_total // world
    nbuf_atoms = n0 if args.workload == "1m" else int(natoms_total * B1 / (0.15 * box[0, 1]))
    sim, kw = build_sim(x, box, local, args.regime, rank=rank, world=world, buffer_size=B1,
                        nbuf_atoms=nbuf_atoms)
    del x
    stream = torch.cuda.ExternalStream(sim.L.obmd_stream(sim.h), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = 1
    sim.run(step, args.equil); step += args.equil          # untimed equilibration from v = 0
    for _ in range(args.warmup):
        sim.run(step, 1); step += 1
    sim.timing(True); sim.timing_reset()
    st0 = sim.fix_state()
    tr0, rd0, in0, de0 = int(st0.ntrials), int(st0.nrounds), int(st0.ninserted), int(st0.ndeleted)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = sim.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    atom_steps = 0
    step_wall = []
    e0.record(stream)
    for _ in range(args.steps):
        tw = time.perf_counter()
        sim.run(step, 1); step += 1
        step_wall.append(time.perf_counter() - tw)   # obmd_run returns after a stream synchronize
        atom_steps += sim.nlocal
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launches - l0
    sampler.stop_flag = True; sampler.join()
    fam = {}
    for name in ("nve", "bin", "sort", "comm", "pre_exchange", "usher", "usher_round0", "usher_rerun",
                 "usher_commit", "usher_allreduce", "usher_eval0", "usher_evalN", "usher_fix", "host_wait_scalars", "host_step", "host_cells", "host_comm", "host_pre_exchange", "host_usher",
                 "host_allreduce", "host_post", "host_final_sync", "pbc", "pair", "post_force"):
        t, nl = sim.timing_get(name)
        fam[name] = {"ms_per_step": t / args.steps, "launches_per_step": nl / args.steps}
    sim.timing(False)
    st = sim.fix_state()
    if world > 1 and os.environ.get("OBMD_BENCH_RANK_TIMERS"):
        print("rank %d: %.3f ms/step local; %s" % (rank, ms / args.steps, {k: round(v["ms_per_step"], 4)
              for k, v in fam.items()}), file=sys.stderr, flush=True)

    # ---- e2e through the C ABI with host buffers (plugin flow) ----
    e2e = None
    if args.e2e_steps > 0:
        e2e = e2e_run(sim, args.e2e_steps, step, barrier)
        step += args.e2e_steps

    # ---- the roofline kernel by itself: in the step it shares the GPU with the USHER rounds (side
    # stream), so its in-step duration measures the overlap, not the kernel.  A few stand-alone launches
    # on the final configuration, L2 flushed in between, CUDA events on the library's stream.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sim.build_cells()
    sim.timing(True); sim.timing_reset()
    n_alone = 5
    for k in range(n_alone):
        flush.fill_(k & 0xff); torch.cuda.synchronize()
        sim.pair_compute(step + k)
    pair_alone_ms = sim.timing_get("pair")[0] / n_alone
    sim.timing(False)
    del flush

    t = torch.tensor([ms, float(atom_steps), e2e["seconds"] if e2e else 0.0,
                      float(e2e["atom_steps"]) if e2e else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = tmax[0].item(); atom_steps = tsum[1].item()
        if e2e:
            e2e["seconds"] = tmax[2].item(); e2e["atom_steps"] = tsum[3].item()
    clocks = sampler.summary()
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        ok = [c for c in allc if c.get("sm_mhz")]
        if ok:
            clocks = dict(min(ok, key=lambda c: c["sm_mhz"]))      # the slowest-clocked rank
            clocks["reasons"] = sorted({r for c in allc for r in c.get("reasons", [])})
            clocks["power_w_max"] = max(c.get("power_w_max", 0.0) for c in ok)
            clocks["ranks"] = world
    if rank == 0:
        hbm, how = peaks()
        pair_ms = fam["pair"]["ms_per_step"]
        achieved = PAIR_BYTES_PER_ATOM * sim.nlocal / (pair_alone_ms * 1e-3) / 1e9 if pair_alone_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "pair_dram_bytes.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": atom_steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "OBMD-DPD fluid, shipped configuration tiled %dx%dx%d = %d atoms (%d per GPU), "
                                   "regime %s (alpha %.1f, nbuf %d, buffer %.2f), T=1, dt=0.001464, rc=1, "
                                   "usher 31.03" % (reps + (natoms_total, n0, args.regime, kw["alpha"],
                                                            int(kw["nbuf"]), B1)),
                       "atoms_per_gpu": n0,
                       "parallelism": ("x-slab decomposition over %d GPUs: NCCL send/recv halo + migration, "
                                       "packed allreduces" % world) if world > 1 else "single GPU",
                       "l2": "per-step working set (%.0f MB) exceeds the 126 MB L2" % (n0 * 168 / 1e6),
                       "equil_steps": args.equil,
                       "inserted": int(st.ninserted), "deleted": int(st.ndeleted),
                       "usher_trials": int(st.ntrials), "usher_rounds": int(st.nrounds)},
            "gpu_launches": int(launches),
            # host wall time of the individual steps on rank 0 (diagnostics: the GPU hosts are noisy
            # virtual machines; `ms_per_step` above is the contract's total / K)
            "ms_per_step_median": round(1e3 * float(np.median(step_wall)), 5),
            "ms_per_step_p90": round(1e3 * float(np.percentile(step_wall, 90)), 5),
            "kernel_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in fam.items()},
            "roofline": {"kernel": "k_pair_row", "bound": "hbm", "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "peak_source": how,
                         "algorithmic_bytes_per_atom": PAIR_BYTES_PER_ATOM,
                         "launch_ms": round(pair_alone_ms, 5), "in_step_ms": round(pair_ms, 5),
                         "note": "launch_ms: %d stand-alone launches on rank 0 after the timed region, L2 flushed "
                                 "between them; in_step_ms: the same kernel inside the step, where it shares the "
                                 "GPU with the USHER rounds on the other stream.  Issue/latency bound in fp64, "
                                 "not HBM bound: DESIGN.md 5.5" % n_alone},
            "clocks": clocks,
        }
        # USHER is latency-bound (a chain of up to nattempt dependent energy evaluations per trial),
        # not bandwidth-bound: report its share of the step and its cost per trial instead of a roofline
        ush = fam["usher"]["ms_per_step"]
        ntr_step = (int(st.ntrials) - tr0) / args.steps
        line["usher"] = {"ms_per_step": round(ush, 5), "share_of_step": round(ush / (ms / args.steps), 4),
                         "eval_ms_per_step": round(fam["usher_eval0"]["ms_per_step"] +
                                                   fam["usher_evalN"]["ms_per_step"], 5),
                         "trials_per_step": round(ntr_step, 1),
                         "rounds_per_step": round((int(st.nrounds) - rd0) / args.steps, 2),
                         "insertions_per_step": round((int(st.ninserted) - in0) / args.steps, 1),
                         "deletions_per_step": round((int(st.ndeleted) - de0) / args.steps, 1),
                         "us_per_trial": round(1e3 * ush / ntr_step, 4) if ntr_step > 0 else None,
                         "note": "rank 0%s; overlapped with the pair kernel on a side stream" %
                                 (" (one of the two edge ranks)" if world > 1 else "")}
        if e2e:
            line["e2e"] = {"value": e2e["atom_steps"] / e2e["seconds"], "unit": UNIT,
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "steps": args.e2e_steps}
        if world == 1 and args.lammps_steps > 0:
            line["e2e_lammps_plugin"] = lammps_plugin_e2e(tuple(args.reps), args.lammps_steps)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        sim.comm_finalize()
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def e2e_run(sim, nsteps, step, barrier):
    """End to end through the C ABI with HOST-owned atom arrays (pinned): every step uploads x, v
    (obmd_upload_xv, 48 B/atom), runs one device step (obmd_run), and reads x, v, f and -- because
    the fix inserts and deletes every step -- tag, type, image back (obmd_download, 84 B/atom),
    so the host copy is current after every step, as a host-resident caller needs it."""
    import torch
    n = sim.nlocal
    cap = int(n * 1.1) + 4096
    pin = lambda *shape, dt=torch.float64: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    buf = {"x": pin(cap, 3), "v": pin(cap, 3), "f": pin(cap, 3), "tag": pin(cap, dt=torch.int32),
           "type": pin(cap, dt=torch.int32), "image": pin(cap, dt=torch.int32)}
    o = sim.download(out=buf)
    h2d = d2h = atom_steps = 0
    barrier()
    t0 = time.perf_counter()
    for s in range(nsteps):
        sim.upload_xv(o["x"], o["v"]); h2d += 48 * len(o["x"])
        sim.run(step + s, 1)
        o = sim.download(out=buf)
        n = len(o["x"]); d2h += 84 * n
        atom_steps += n
    barrier()
    sec = time.perf_counter() - t0
    return {"seconds": sec, "atom_steps": atom_steps, "h2d": h2d // nsteps, "d2h": d2h // nsteps}


# ------------------------------------------------------------------ reference arm / CPU baseline

DECK = """units lj
boundary f p p
atom_style atomic
comm_modify vel yes
newton on
read_data {data}
region leftB block 0.0 {xlB:.17g} 0.0 {yhi:.17g} 0.0 {zhi:.17g}
region rightB block {xrB:.17g} {xhi:.17g} 0.0 {yhi:.17g} 0.0 {zhi:.17g}
region leftshear block 0.0 0.0 0.0 0.0 0.0 0.0
region rightshear block 0.0 0.0 0.0 0.0 0.0 0.0
pair_style dpd 1.0 1.0 2616
pair_coeff * * 209.6 4.5 1.0
neighbor 0.4 bin
neigh_modify delay 0 every 1
timestep 0.001464
fix 1 all nve
fix 2 all obmd 1 1 6111 188.0 0 0 0 0 {alpha} 0.005 {nbuf} region1 leftB region2 rightB region3 leftshear region4 rightshear region5 leftB region6 rightB buffersize {B:.17g} gfac 0.25 stepparallel 0 stepperp 1 maxattempt 1 usher 1 31.03 1.0 0.02 10000 1.5 1.0 40 charged 0
thermo_style custom step atoms
thermo {nsteps}
run {nsteps}
"""


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
    return p if os.path.exists(p) else None


def run_reference(reps, nsteps, regime="R1", plugin=None):
    """Run the unmodified reference (oracle/_ref/lmp_serial, 1 core, MPI stubs) on the tiled deck;
    returns (atom_steps, loop_seconds)."""
    x, box = tiled(reps)
    n = len(x)
    wd = tempfile.mkdtemp(prefix="obmd_ref_")
    data = os.path.join(wd, "tiled.data")
    with open(data, "w") as fh:
        fh.write("tiled OBMD_DPD\n\n%d atoms\n1 atom types\n\n0.0 %.17g xlo xhi\n0.0 %.17g ylo yhi\n"
                 "0.0 %.17g zlo zhi\n\nMasses\n\n1 1\n\nAtoms\n\n" % (n, box[0, 1], box[1, 1], box[2, 1]))
        ids = np.arange(1, n + 1)
        np.savetxt(fh, np.column_stack([ids, np.ones(n), x]), fmt="%d %d %.6f %.6f %.6f")
    B = 0.15 * box[0, 1]
    alpha, nbuf = (1.0, int(0.15 * n)) if regime == "R1" else (0.7, int(0.7 * 0.15 * n))
    deck = ("plugin load %s\n" % plugin if plugin else "") + DECK.format(data=data, xlB=B, xrB=box[0, 1] - B, xhi=box[0, 1], yhi=box[1, 1], zhi=box[2, 1],
                       alpha=alpha, nbuf=nbuf, B=B, nsteps=nsteps)
    inp = os.path.join(wd, "in.ref")
    open(inp, "w").write(deck)
    out = os.path.join(wd, "out.ref")
    with open(out, "w") as fh:     # the reference prints ~40 debug lines per step
        subprocess.check_call([ref_binary(), "-in", inp, "-log", "none"], cwd=wd, stdout=fh,
                              stderr=subprocess.STDOUT)
    loop = None
    atoms = []
    for ln in open(out, errors="replace"):
        w = ln.split()
        if ln.startswith("Loop time of"):
            loop = float(w[3])
        elif len(w) == 2 and w[0].isdigit() and w[1].isdigit():
            atoms.append(int(w[1]))
    for f in (data, inp, out):
        os.remove(f)
    navg = 0.5 * (atoms[0] + atoms[-1]) if atoms else n
    return navg * nsteps, loop


def lammps_plugin_e2e(reps, nsteps):
    """The real drop-in: the unmodified reference LAMMPS binary with `plugin load
    libobmd_b200_lammps.so` (fix obmd + pair_style dpd on the GPU, FixNVE/pbc/comm on the host)."""
    plug = os.path.join(ROOT, "obmd-lammps-extension_b200", "libobmd_b200_lammps.so")
    if ref_binary() is None or not os.path.exists(plug):
        return {"value": None, "unit": UNIT, "note": "reference binary or plugin not built"}
    a, loop = run_reference(reps, nsteps, plugin=plug)
    return {"value": a / loop, "unit": UNIT, "steps": nsteps,
            "note": "oracle/_ref/lmp_serial + plugin, LAMMPS loop time %.3f s; host FixNVE prints debug "
                    "lines every step (SURVEY Q8)" % loop}


def ref_cost_per_step(n, area_tiles):
    """Rough serial cost model of the reference: ~2.1 us/atom-step for
    neigh+pair+fix scans, plus brute-force USHER: ~130 trials per shipped cross-section per step,
    ~5 iterations each, ~3.5 ns per atom per iteration."""
    return 2.1e-6 * n + 130 * area_tiles * 5 * 3.5e-9 * n


def pick_sample(total_steps, budget_s):
    for reps in ((4, 5, 4), (4, 2, 2), (4, 1, 2), (4, 1, 1), (2, 1, 1), (1, 1, 1)):
        n = 12637 * reps[0] * reps[1] * reps[2]
        if (ref_cost_per_step(n, reps[1] * reps[2]) + 0.0) * total_steps + 4e-6 * n <= budget_s:
            return reps
    return (1, 1, 1)


def cpu_baseline(budget_s=25.0, steps=4, reps=None):
    if ref_binary() is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/lmp_serial missing (run oracle/build_ref.sh)"}
    reps = reps or pick_sample(steps, budget_s)
    a, loop = run_reference(reps, steps)
    n = 12637 * reps[0] * reps[1] * reps[2]
    return {"value": a / loop, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "%d steps of the R1 deck tiled %dx%dx%d (%d atoms), unmodified reference "
                      "(LAMMPS 7Feb2024 + src/OBMD, g++ -O3, MPI stubs, 1 core), LAMMPS loop time %.2f s"
                      % (steps, reps[0], reps[1], reps[2], n, loop)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    reps = pick_sample(total, args.ref_budget)
    n = 12637 * reps[0] * reps[1] * reps[2]
    if ref_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lmp_serial not built"}))
        return
    # warm-up and timed steps are separate runs of the binary: LAMMPS' loop time excludes setup
    if args.warmup > 0:
        run_reference(reps, args.warmup)
    a, loop = run_reference(reps, args.steps)
    val = a / loop
    x, box = tiled(tuple(args.reps))
    scaling = "weak" if args.workload == "1m" else "strong"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * loop / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "OBMD-DPD fluid, shipped configuration tiled %dx%dx%d = %d atoms, regime R1; "
                                   "reference timed on a bounded sample: tiling %dx%dx%d = %d atoms"
                                   % (tuple(args.reps) + (len(x),) + reps + (n,))},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                             "sample": "%d timed steps, tiling %dx%dx%d (%d atoms), lmp_serial 1 core"
                                       % (args.steps, reps[0], reps[1], reps[2], n)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reps", type=int, nargs=3, default=[4, 5, 4], help="tiling of the shipped box")
    ap.add_argument("--workload", default="1m", choices=["1m", "4m", "16m"],
                    help="1m: 1 010 960 atoms per GPU (weak scaling); 4m/16m: fixed total (strong scaling)")
    ap.add_argument("--regime", default="R1", choices=["R0", "R1"])
    ap.add_argument("--lammps-steps", type=int, default=0,
                    help="also time the reference LAMMPS binary + plugin (drop-in e2e) for this many steps")
    ap.add_argument("--equil", type=int, default=300, help="untimed equilibration steps from v = 0")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-budget", type=float, default=150.0)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
