#!/usr/bin/env python
"""bench.py -- OBMD-DPD atom-steps/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path

A "step" is one MD timestep of the whole hot path (FixNVE halves, cell build, fix obmd
pre_exchange incl. batched USHER, PairDPD, fix obmd post_force) over every atom.

Start state of the named workloads: the shipped configuration (uniform DPD fluid, rho = 3) tiled by
integer translations, Maxwell-Boltzmann velocities at T = 1, then --equil (default 2000) untimed
steps of the full open-boundary step so that the density / velocity profiles at the open faces and
the trial / insertion / deletion rates are the steady ones before the timed region starts.
`like_for_like` is the one workload BOTH arms can run from the SAME state: the shipped system after
2000 steps of the UNMODIFIED reference in regime R1 (obmd-lammps-extension_b200/data/equil_r1.npz,
written by tests/golden/make_golden.py r1long: positions AND velocities), tiled across the periodic
cross-section only -- an exact steady state, which the reference arm is timed on as well.

N = 1: BASELINE config 2, "OBMD-DPD fluid at equilibrium, 1M particles" = 4x5x4 tiles, regime R1
(alpha 1.0, nbuf = int(0.15 N)).  Secondary objects in the same line: `r0` (the shipped parameters,
alpha 0.7: buffers only drain, no USHER) and `strong_16m` (BASELINE config 5 on this GPU count).
N > 1: BASELINE config 5, 16 M particles (20x8x8 tiles) slab-decomposed along x, strong scaling,
buffer 0.15 Lx at every N; `weak_1m_per_gpu` is the secondary object.
value = sum over steps of the atom count / device time (CUDA events on the library's stream, max
over ranks).  e2e = the same metric through the C ABI with HOST-owned atom arrays: every step uploads
x,v from pinned memory (obmd_upload_xv), runs one device step (obmd_run) and reads x,v,f,tag,type,
image back into pinned memory (obmd_download).
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

METRIC = "OBMD-DPD atom-steps/sec"
UNIT = "atom-steps/s"
PAIR_BYTES_PER_ATOM = 80.0   # x,v,tag,type in + f out
NVE_BYTES_PER_ATOM = 192.0   # initial 72 r + 48 w, final 48 r + 24 w
CELLS_BYTES_PER_ATOM = 236.0  # bin 24 + 8, perm 12, gather 64 r + 64 w + 64 w of records
POST_BYTES_PER_BUFFER_ATOM = 100.0
USHER = (1, 31.03, 1.0, 0.02, 1e4, 1.5, 1.0, 40)
TILE_LX = 33.59462486002239
N0_TILE = 12637              # atoms of the shipped data file: nbuf is defined from it (README.UserGuide.md:190-233)
TILINGS = {"1m": (4, 5, 4), "4m": (5, 8, 8), "16m": (20, 8, 8)}


SHEAR_NOTE = ("BASELINE config 3: 4 M particles under a prescribed oscillatory tangential momentum flux, pxy(t) = "
              "0.5 sin(2 pi 0.1 t) set through obmd_fix_set_flux before every step (what the plugin does for v_ "
              "arguments), shear regions a quarter of the buffer long; 1 and 2 GPUs")
DATA = os.path.join(ROOT, "obmd-lammps-extension_b200", "data")


def tiled(reps, snapshot=False):
    """snapshot=False: the shipped configuration (examples/OBMD_DPD/dpd_8map_obmd.data) tiled reps times,
    Maxwell-Boltzmann velocities at T = 1 (seeded).  snapshot=True: the reference's own R1 steady state
    (positions + velocities); only meaningful with reps[0] == 1 (the open faces are part of it)."""
    if snapshot:
        d = np.load(os.path.join(DATA, "equil_r1.npz"))
        x, v, box = d["x"].copy(), d["v"].copy(), d["box"].copy()
    else:
        d = np.load(os.path.join(DATA, "dpd_8map_obmd.npz"))
        x, v, box = d["x"].copy(), None, d["box"].copy()
    L = box[:, 1] - box[:, 0]
    shifts = [np.array([i, j, k]) * L for i in range(reps[0]) for j in range(reps[1]) for k in range(reps[2])]
    X = np.concatenate([x + s for s in shifts])
    if v is None:
        V = np.random.default_rng(4928459).normal(0.0, 1.0, X.shape)
        V -= V.mean(0)
    else:
        V = np.concatenate([v] * len(shifts))
    B = np.array([[0.0, L[0] * reps[0]], [0.0, L[1] * reps[1]], [0.0, L[2] * reps[2]]])
    return X, V, B


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


# ------------------------------------------------------------------ our arm

def regime_kw(regime, ntiles, buffer_frac=1.0):
    """alpha and nbuf of the regime; nbuf = the nominal buffer population 0.15 N (README.UserGuide.md:190-233),
    scaled when the buffer is shorter than 0.15 Lx (weak scaling keeps the N = 1 physical length)."""
    n_nominal = N0_TILE * ntiles * buffer_frac
    if regime == "R1":
        return dict(alpha=1.0, nbuf=float(int(0.15 * n_nominal)))
    return dict(alpha=0.7, nbuf=float(int(0.7 * 0.15 * n_nominal)))


def fix_args(ob, box, buffer_size, alpha, nbuf, shear_size=0.0):
    """`fix obmd 1 1 6111 188.0 0 0 0 0 alpha 0.005 nbuf ... usher 1 31.03 1.0 0.02 10000 1.5 1.0 40`
    (README.UserGuide.md:190-233) with regions built from the exact box bounds."""
    a = ob.FixArgs()
    a.ntype, a.nfreq, a.seed = 1, 1, 6111
    a.pxx, a.pxy, a.pxz, a.dpxx, a.freq = 188.0, 0.0, 0.0, 0.0, 0.0
    a.alpha, a.tau, a.nbuf = alpha, 0.005, nbuf
    a.buffer_size, a.g_fac = buffer_size, 0.25
    a.step_parallel, a.step_perp, a.maxattempt = 0, 1, 1
    (a.usherflag, a.etarget, a.ds0, a.dtheta0, a.uovlp, a.dsovlp, a.eps, a.nattempt) = USHER
    for k, (lo, hi) in enumerate(ob.default_regions(box[:, 0], box[:, 1], buffer_size, shear_size)):
        for d in range(3):
            a.region[k].lo[d] = lo[d]
            a.region[k].hi[d] = hi[d]
    return a


def build_sim(reps, device, regime, rank=0, world=1, buffer_size=None, snapshot=False, shear=False):
    """Context for this rank: pair_style dpd 1.0 1.0 2616, pair_coeff * * 209.6 4.5 1.0, dt 0.001464.
    world > 1: x-slab decomposition (obmd_comm_init), every rank uploads only the atoms inside its
    slab; the NCCL unique id travels through torch.distributed."""
    import obmd_b200 as ob
    x, v, box = tiled(reps, snapshot)
    ntiles = reps[0] * reps[1] * reps[2]
    B = buffer_size if buffer_size is not None else 0.15 * box[0, 1]
    kw = regime_kw(regime, ntiles, B / (0.15 * box[0, 1]))
    sim = ob.ObmdB200(device)
    sim.set_box(box[:, 0], box[:, 1])
    sim.set_masses([1.0])
    sim.set_timestep(0.001464)
    sim.pair_style_dpd(1.0, 1.0, 2616)
    sim.pair_coeff("*", "*", 209.6, 4.5, 1.0)
    sim.pair_init(1.0)
    sim.fix_obmd(fix_args(ob, box, B, kw["alpha"], kw["nbuf"], 0.25 * B if shear else 0.0))
    n = len(x)
    tag = np.arange(1, n + 1, dtype=np.int32)
    if world > 1:
        import torch.distributed as dist
        uid = [ob.ObmdB200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.comm_init(rank, world, uid[0])
        lo, hi = sim.comm_slab()
        mine = (x[:, 0] >= lo) & ((x[:, 0] < hi) | ((rank == world - 1) & (x[:, 0] <= box[0, 1])))
        x, v, tag = x[mine], v[mine], tag[mine]
    m = len(x)
    sim.upload(x, v, np.ones(m, np.int32), tag, capacity=int(1.15 * m) + 65536)
    sim.setup(0)
    return sim, dict(kw, buffer=B, natoms=n, box=box, local=m)


FAMILIES = ("nve", "bin", "sort", "comm", "pre_exchange", "usher", "usher_eval0", "usher_fix", "usher_commit",
            "usher_allreduce", "pbc", "pair", "post_force", "host_wait_scalars", "host_step", "host_cells",
            "host_comm", "host_pre_exchange", "host_usher", "host_post", "host_final_sync")


def timed_run(sim, stream, barrier, steps, warmup, equil, sampler=None, per_step=None):
    """per_step(step): host work before every step (the plugin evaluates v_ variables of `fix obmd` there
    and calls obmd_fix_set_flux; fix_obmd_merged.cpp:563-572)."""
    import torch
    step = 1
    if equil > 0:
        if per_step is None:
            sim.run(step, equil); step += equil
        else:
            for _ in range(equil):
                per_step(step); sim.run(step, 1); step += 1
    for _ in range(warmup):
        if per_step:
            per_step(step)
        sim.run(step, 1); step += 1
    sim.timing(True); sim.timing_reset()
    st0 = sim.fix_state()
    if sampler:
        sampler.start()
    barrier()
    l0 = sim.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    atom_steps, wall = 0, []
    e0.record(stream)
    for _ in range(steps):
        tw = time.perf_counter()
        if per_step:
            per_step(step)
        sim.run(step, 1); step += 1
        wall.append(time.perf_counter() - tw)   # obmd_run returns after a stream synchronize
        atom_steps += sim.nlocal
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if sampler:
        sampler.stop_flag = True; sampler.join()
    fam = {}
    for name in FAMILIES:
        t, nl = sim.timing_get(name)
        fam[name] = {"ms_per_step": t / steps, "launches_per_step": nl / steps}
    sim.timing(False)
    st1 = sim.fix_state()
    d = lambda k: (int(getattr(st1, k)) - int(getattr(st0, k))) / steps
    return {"ms": ms, "atom_steps": atom_steps, "fam": fam, "launches": sim.launches - l0, "step": step,
            "wall": wall, "trials": d("ntrials"), "rounds": d("nrounds"), "ins": d("ninserted"),
            "dele": d("ndeleted")}


def reduce_over_ranks(ms, atom_steps, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return ms, atom_steps
    t = torch.tensor([ms, float(atom_steps)], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return tmax[0].item(), tsum[1].item()


def secondary(reps, local, regime, rank, world, steps, warmup, equil, scaling, buffer_size=None, snapshot=False,
              shear=False):
    """One more workload measured the same way (no e2e, no roofline): a compact object for the line.
    shear: BASELINE config 3 -- shear regions of a quarter of the buffer length and a tangential momentum
    flux pxy(t) = 0.5 sin(2 pi 0.1 t) prescribed through obmd_fix_set_flux before every step."""
    import torch
    import torch.distributed as dist
    sim, info = build_sim(reps, local, regime, rank=rank, world=world, buffer_size=buffer_size, snapshot=snapshot,
                          shear=shear)
    per_step = None
    if shear:
        def per_step(step):
            pxy = 0.5 * np.sin(2.0 * np.pi * 0.1 * step * 0.001464)
            sim.set_flux(188.0, pxy, 0.0, 0.0, 0.0, info["alpha"], 0.005, info["nbuf"])
    stream = torch.cuda.ExternalStream(sim.L.obmd_stream(sim.h), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    r = timed_run(sim, stream, barrier, steps, warmup, equil, per_step=per_step)
    ms, atom_steps = reduce_over_ranks(r["ms"], r["atom_steps"], world)
    out = {"value": atom_steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "n_gpus": world,
           "scaling": scaling, "steps": steps, "equil_steps": equil,
           "workload": "tiling %dx%dx%d = %d atoms, regime %s, buffer %.2f" % (reps + (info["natoms"], regime, info["buffer"])),
           "trials_per_step": round(r["trials"], 1), "insertions_per_step": round(r["ins"], 1),
           "deletions_per_step": round(r["dele"], 1),
           "kernel_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in r["fam"].items()
                                  if not k.startswith("host") and v["ms_per_step"] > 0}}
    chk = slab_check(sim, info, world) if world > 1 else None
    if chk:
        out["slab_check"] = chk
    if world > 1:
        sim.comm_finalize()
    sim.close()
    return out


def slab_check(sim, info, world):
    """Invariants of a slab run, gathered over the ranks: tags unique, atom count = N0 - deleted +
    inserted (the fix's counters are global), every atom inside its rank's slab (+- the distance it
    can move in one step)."""
    import torch
    import torch.distributed as dist
    o = sim.download(("x", "tag"))
    st = sim.fix_state()
    lo, hi = sim.comm_slab()
    n = len(o["tag"])
    inside = bool(np.all((o["x"][:, 0] >= lo - 0.5) & (o["x"][:, 0] <= hi + 0.5))) if n else True
    tags = torch.tensor(o["tag"].astype(np.int64), device="cuda")
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device="cuda"))
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    mine = torch.full((pad,), -1, dtype=torch.int64, device="cuda")
    mine[:n] = tags
    buf = [torch.zeros(pad, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(buf, mine)
    allt = torch.cat([b[:s] for b, s in zip(buf, sizes)])
    flags = torch.tensor([1 if inside else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    total = int(allt.numel())
    return {"atoms": total, "atoms_per_rank": sizes, "tags_unique": bool(torch.unique(allt).numel() == total),
            "count_balance_ok": total == info["natoms"] - int(st.ndeleted) + int(st.ninserted),
            "inserted": int(st.ninserted), "deleted": int(st.ndeleted), "atoms_inside_own_slab": bool(flags.item())}


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
        # has host round trips, and rank threads migrating onto each other's cores showed up as
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
    workload = args.workload or ("1m" if world == 1 else "16m")
    if workload == "1m":       # one 4x5x4 tiling per GPU along the open axis (weak scaling at N > 1)
        r1 = tuple(args.reps)
        reps, scaling = (r1[0] * world, r1[1], r1[2]), "weak"
        buffer_size = 0.15 * TILE_LX * r1[0]          # the N = 1 physical buffer length
    else:                      # BASELINE configs 3-5: fixed total size (strong scaling), buffer 0.15 Lx at every N
        reps, scaling = TILINGS[workload], "strong"
        buffer_size = 0.15 * TILE_LX * reps[0]
    sim, info = build_sim(reps, local, args.regime, rank=rank, world=world, buffer_size=buffer_size)
    natoms_total, n0 = info["natoms"], info["natoms"] // world
    stream = torch.cuda.ExternalStream(sim.L.obmd_stream(sim.h), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    r = timed_run(sim, stream, barrier, args.steps, args.warmup, args.equil, sampler)
    fam, step = r["fam"], r["step"]
    if world > 1 and os.environ.get("OBMD_BENCH_RANK_TIMERS"):
        print("rank %d: %.3f ms/step local; %s" % (rank, r["ms"] / args.steps, {k: round(v["ms_per_step"], 4)
              for k, v in fam.items()}), file=sys.stderr, flush=True)

    # ---- e2e through the C ABI with host buffers (plugin flow) ----
    e2e = None
    if args.e2e_steps > 0:
        e2e = e2e_run(sim, args.e2e_steps, step, barrier)
        step += args.e2e_steps

    # ---- the roofline kernel by itself: in the step it shares the GPU with USHER (side stream), so its
    # in-step duration measures the overlap, not the kernel.  A few stand-alone launches on the final
    # configuration, L2 flushed in between, CUDA events on the library's stream.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sim.build_cells()
    if world > 1:
        sim.exchange()
    # no pre_exchange here: inside obmd_run the pair kernel is launched on the freshly sorted atoms,
    # before USHER appends this step's insertions (those get their forces from k_pair_add_new)
    sim.timing(True); sim.timing_reset()
    n_alone = 5
    for k in range(n_alone):
        flush.fill_(k & 0xff); torch.cuda.synchronize()
        sim.pair_compute(step + k)
    pair_alone_ms = sim.timing_get("pair")[0] / n_alone
    sim.timing(False)
    del flush
    nlocal_final = sim.nlocal

    ms, atom_steps = reduce_over_ranks(r["ms"], r["atom_steps"], world)
    if e2e and world > 1:
        e2e["seconds"], e2e["atom_steps"] = reduce_over_ranks(e2e["seconds"], e2e["atom_steps"], world)
    clocks = sampler.summary()
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        ok = [c for c in allc if c.get("sm_mhz")]
        if ok:
            clocks = dict(min(ok, key=lambda c: c["sm_mhz"]))      # the slowest-clocked rank
            clocks["reasons"] = sorted({r_ for c in allc for r_ in c.get("reasons", [])})
            clocks["power_w_max"] = max(c.get("power_w_max", 0.0) for c in ok)
            clocks["ranks"] = world
    chk = slab_check(sim, info, world) if world > 1 else None
    if world > 1:
        sim.comm_finalize()
    sim.close()

    # ---- secondary workloads (same protocol, fewer steps) ----
    extra = {}
    sec_steps = max(5, min(args.steps, args.secondary_steps))
    if not args.no_secondary:
        if world == 1:
            extra["r0"] = secondary(reps, local, "R0", rank, world, sec_steps, args.warmup, min(args.equil, 300),
                                    scaling, buffer_size)
            if rank == 0:
                extra["r0"]["note"] = ("regime R0: the parameters exactly as shipped (alpha 0.7); the buffers only "
                                       "drain, no USHER -- the step is cell build + PairDPD + force distribution + deletion")
            extra["r1_equil300"] = secondary(reps, local, "R1", rank, world, max(sec_steps, 100), args.warmup, 300,
                                             scaling, buffer_size)
            extra["r1_equil300"]["note"] = ("the primary workload under the round-1 protocol (300 equilibration steps "
                                            "instead of 2000: the USHER trial count is still climbing, ~4.5 k per step "
                                            "instead of ~13 k), kept so that round-over-round numbers are comparable")
            extra["strong_16m"] = secondary(TILINGS["16m"], local, "R1", rank, world, sec_steps, args.warmup,
                                            args.equil, "strong")
            extra["strong_16m"]["note"] = ("BASELINE config 5 on ONE GPU, same equilibration as the N > 1 runs: the N = 1 point of the strong-scaling series "
                                           "`bench.py --gpus N` reports as its primary value for N > 1")
            extra["shear_4m"] = secondary(TILINGS["4m"], local, "R1", rank, world, sec_steps, args.warmup,
                                          min(args.equil, 300), "strong", shear=True)
            extra["shear_4m"]["note"] = SHEAR_NOTE
            lreps = tuple(args.like_reps)
            extra["like_for_like"] = secondary(lreps, local, "R1", rank, world, args.steps, args.warmup, 0, "n/a",
                                               None, snapshot=True)
            extra["like_for_like"]["note"] = ("the workload `bench.py --impl reference` times: the reference's own R1 "
                                              "steady state (positions + velocities after 2000 of its steps) tiled "
                                              "%dx%dx%d across the periodic cross-section; same deck, same start state, "
                                              "no equilibration on either side.  Far too small to fill a B200: it is "
                                              "here so that one ratio is between identical configurations" % lreps)
        elif workload != "1m":
            if world == 2:
                extra["shear_4m"] = secondary(TILINGS["4m"], local, "R1", rank, world, sec_steps, args.warmup,
                                              min(args.equil, 300), "strong", shear=True)
                extra["shear_4m"]["note"] = SHEAR_NOTE
            r1 = tuple(args.reps)
            extra["weak_1m_per_gpu"] = secondary((r1[0] * world, r1[1], r1[2]), local, "R1", rank, world, sec_steps,
                                                 args.warmup, min(args.equil, 300), "weak", 0.15 * TILE_LX * r1[0])
            # the N = 1 point of this strong-scaling series, measured in the same run on rank 0's GPU (the
            # other ranks wait): the driver's own N = 1 run times BASELINE config 2 (1 M atoms), not this workload
            if rank == 0:
                extra["n1_same_workload"] = secondary(reps, local, args.regime, 0, 1, sec_steps, args.warmup,
                                                      args.equil, "strong")
                extra["n1_same_workload"]["note"] = ("the primary workload of this line on ONE GPU (rank 0's), same "
                                                     "equilibration: the denominator of the strong-scaling efficiency")
            dist.barrier()
        else:
            extra["strong_16m"] = secondary(TILINGS["16m"], local, "R1", rank, world, sec_steps, args.warmup,
                                            min(args.equil, 300), "strong")

    if rank == 0:
        hbm, how = peaks()
        pair_ms = fam["pair"]["ms_per_step"]
        achieved = PAIR_BYTES_PER_ATOM * nlocal_final / (pair_alone_ms * 1e-3) / 1e9 if pair_alone_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "pair_dram_bytes.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        kw = info
        line = {
            "metric": METRIC, "value": atom_steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(reps, natoms_total, args.regime, kw),
                       "atoms_per_gpu": n0,
                       "parallelism": ("x-slab decomposition over %d GPUs: NCCL send/recv halo + migration, "
                                       "packed allreduces" % world) if world > 1 else "single GPU",
                       "l2": "per-step working set (%.0f MB per GPU) exceeds the 126 MB L2" % (n0 * 168 / 1e6),
                       "start_state": "shipped configuration tiled, Maxwell-Boltzmann velocities T = 1, then "
                                      "equil_steps untimed steps of the full step",
                       "equil_steps": args.equil},
            "gpu_launches": int(r["launches"]),
            # host wall time of the individual steps on rank 0 (diagnostics: the GPU hosts are noisy
            # virtual machines; `ms_per_step` above is the contract's total / K)
            "ms_per_step_median": round(1e3 * float(np.median(r["wall"])), 5),
            "ms_per_step_p90": round(1e3 * float(np.percentile(r["wall"], 90)), 5),
            "kernel_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in fam.items()},
            "roofline": {"kernel": "k_pair_brick", "bound": "hbm", "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "peak_source": how,
                         "algorithmic_bytes_per_atom": PAIR_BYTES_PER_ATOM,
                         "launch_ms": round(pair_alone_ms, 5), "in_step_ms": round(pair_ms, 5),
                         "note": "launch_ms: %d stand-alone launches on rank 0 after the timed region, L2 flushed "
                                 "between them; in_step_ms: the pair family inside the step, where it shares the "
                                 "GPU with USHER on the other stream.  Issue/latency bound in fp64, "
                                 "not HBM bound: DESIGN.md 5.5" % n_alone},
            "clocks": clocks,
        }
        # the streaming kernel families of the step, from the sync-free CUDA-event timers inside the timed
        # region (rank 0): algorithmic bytes per step / time, against the same measured peak
        nat = atom_steps / args.steps / world
        nbuf_atoms = nat * min(1.0, 2.0 * kw["buffer"] / kw["box"][0, 1] * (world if world <= 2 else world / 2.0)) \
            if world > 1 else nat * 2.0 * kw["buffer"] / kw["box"][0, 1]
        fams = []
        for name, key, kernels, byts, units in (
                ("nve", "nve", "k_nve_initial + k_nve_final", NVE_BYTES_PER_ATOM, nat),
                ("cell build", None, "k_bin, k_scan_*, k_perm, k_cell_order, k_gather", CELLS_BYTES_PER_ATOM, nat),
                ("post_force", "post_force", "k_weight_sums, k_apply_forces",
                 POST_BYTES_PER_BUFFER_ATOM, nbuf_atoms)):
            t = fam[key]["ms_per_step"] if key else fam["bin"]["ms_per_step"] + fam["sort"]["ms_per_step"]
            if t > 0:
                gbs = byts * units / (t * 1e-3) / 1e9
                fams.append({"family": name, "kernels": kernels, "ms_per_step": round(t, 5), "achieved": round(gbs, 1),
                             "unit": "GB/s", "frac": round(gbs / hbm, 4), "algorithmic_bytes_per_unit": byts,
                             "units_per_step": int(units)})
        line["rooflines"] = fams
        # USHER is latency-bound (a chain of up to nattempt dependent energy evaluations per trial),
        # not bandwidth-bound: report its share of the step and its cost per trial instead of a roofline
        ush = fam["usher"]["ms_per_step"]
        line["usher"] = {"ms_per_step": round(ush, 5), "share_of_step": round(ush / (ms / args.steps), 4),
                         "round0_ms_per_step": round(fam["usher_eval0"]["ms_per_step"], 5),
                         "rerun_rounds_ms_per_step": round(fam["usher_fix"]["ms_per_step"], 5),
                         "trials_per_step": round(r["trials"], 1), "rounds_per_step": round(r["rounds"], 2),
                         "insertions_per_step": round(r["ins"], 1), "deletions_per_step": round(r["dele"], 1),
                         "us_per_trial": round(1e3 * ush / r["trials"], 4) if r["trials"] > 0 else None,
                         "note": "rank 0%s; overlapped with the pair kernel on a side stream" %
                                 (" (one of the two edge ranks)" if world > 1 else "")}
        if e2e:
            line["e2e"] = {"value": e2e["atom_steps"] / e2e["seconds"], "unit": UNIT,
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "steps": args.e2e_steps}
        if chk:
            line["slab_check"] = chk
        line.update(extra)
        if world == 1 and args.lammps_steps > 0:
            line["e2e_lammps_plugin"] = lammps_plugin_e2e(tuple(args.reps), args.lammps_steps)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(budget_s=args.cpu_budget, steps=10, reps=tuple(args.like_reps))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def workload_name(reps, natoms, regime, kw):
    return ("OBMD-DPD fluid, shipped configuration tiled %dx%dx%d = %d atoms, "
            "regime %s (alpha %.1f, nbuf %d, buffer %.2f), T=1, dt=0.001464, rc=1, usher 31.03"
            % (tuple(reps) + (natoms, regime, kw["alpha"], int(kw["nbuf"]), kw["buffer"])))


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
atom_modify sort 0 0
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
{run_cmd} {nsteps}
"""


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
    return p if os.path.exists(p) else None


def run_reference(reps, nsteps, regime="R1", plugin=None, snapshot=True, run_cmd="run"):
    """Run the unmodified reference (oracle/_ref/lmp_serial, 1 core, MPI stubs) on a tiled start state
    with velocities (snapshot=True: its own R1 steady state, as `like_for_like` in our arm); returns
    (atom_steps, loop_seconds, usher_trials)."""
    x, v, box = tiled(reps, snapshot)
    n = len(x)
    wd = tempfile.mkdtemp(prefix="obmd_ref_")
    data = os.path.join(wd, "tiled.data")
    ids = np.arange(1, n + 1)
    with open(data, "w") as fh:
        fh.write("tiled OBMD_DPD R1 snapshot\n\n%d atoms\n1 atom types\n\n0.0 %.17g xlo xhi\n0.0 %.17g ylo yhi\n"
                 "0.0 %.17g zlo zhi\n\nMasses\n\n1 1\n\nAtoms\n\n" % (n, box[0, 1], box[1, 1], box[2, 1]))
        np.savetxt(fh, np.column_stack([ids, np.ones(n), x]), fmt="%d %d %.12g %.12g %.12g")
        fh.write("\nVelocities\n\n")
        np.savetxt(fh, np.column_stack([ids, v]), fmt="%d %.12g %.12g %.12g")
    B = 0.15 * box[0, 1]
    kw = regime_kw(regime, reps[0] * reps[1] * reps[2])
    deck = ("plugin load %s\n" % plugin if plugin else "") + DECK.format(
        data=data, xlB=B, xrB=box[0, 1] - B, xhi=box[0, 1], yhi=box[1, 1], zhi=box[2, 1],
        alpha=kw["alpha"], nbuf=int(kw["nbuf"]), B=B, nsteps=nsteps, run_cmd=run_cmd)
    inp = os.path.join(wd, "in.ref")
    open(inp, "w").write(deck)
    out = os.path.join(wd, "out.ref")
    with open(out, "w") as fh:     # the reference prints ~40 debug lines per step
        subprocess.check_call([ref_binary(), "-in", inp, "-log", "none"], cwd=wd, stdout=fh,
                              stderr=subprocess.STDOUT)
    loop, atoms, ntrials = None, [], 0
    for ln in open(out, errors="replace"):
        w = ln.split()
        if ln.startswith("Loop time of"):
            loop = float(w[3])
        elif ln.startswith("USHER"):
            ntrials += 1
        elif len(w) == 2 and w[0].isdigit() and w[1].isdigit():
            atoms.append(int(w[1]))
    for f in (data, inp, out):
        os.remove(f)
    navg = 0.5 * (atoms[0] + atoms[-1]) if atoms else n
    return navg * nsteps, loop, ntrials


def lammps_plugin_e2e(reps, nsteps):
    """The real drop-in: the unmodified reference LAMMPS binary with `plugin load
    libobmd_b200_lammps.so` and `run_obmd N` (stage B: fix nve + fix obmd + pair_style dpd device-resident,
    atoms back on the host only on output steps and at the end)."""
    plug = os.path.join(ROOT, "obmd-lammps-extension_b200", "libobmd_b200_lammps.so")
    if ref_binary() is None or not os.path.exists(plug):
        return {"value": None, "unit": UNIT, "note": "reference binary or plugin not built"}
    a, loop, _ = run_reference(reps, nsteps, plugin=plug, snapshot=False, run_cmd="run_obmd")
    return {"value": a / loop, "unit": UNIT, "steps": nsteps,
            "note": "oracle/_ref/lmp_serial + plugin, `run_obmd` (device-resident), LAMMPS loop time %.3f s, "
                    "from the tiled shipped configuration with thermal velocities (no equilibration)" % loop}


def ref_cost_per_step(reps):
    """Serial cost model of the reference in steady exchange, fitted to its own R1 run of the shipped
    deck (42.5 k atom-steps/s at 12 012 atoms, 91.7 % of it in the fix): ~2.1 us per atom for
    neighbour build + pair + scans, plus the brute-force USHER energy: ~130 trials per shipped
    cross-section per step, each ~1.65e-7 s per owned atom (up to 40 iterations over ALL atoms)."""
    n = 12008 * reps[0] * reps[1] * reps[2]
    return 2.1e-6 * n + 130 * reps[1] * reps[2] * 1.65e-7 * n


def pick_sample(total_steps, budget_s):
    """Largest tiling of the reference's R1 snapshot across the periodic cross-section whose run fits the budget."""
    for reps in ((1, 5, 4), (1, 3, 3), (1, 2, 3), (1, 2, 2), (1, 1, 2), (1, 1, 1)):
        n = 12008 * reps[0] * reps[1] * reps[2]
        if ref_cost_per_step(reps) * total_steps + 6e-6 * n <= budget_s:
            return reps
    return (1, 1, 1)


def sample_text(reps, steps, loop, ntrials):
    n = 12008 * reps[0] * reps[1] * reps[2]
    full = ref_cost_per_step((4, 5, 4))
    return ("%d steps of the same deck from the reference's own R1 steady state (positions + velocities after "
            "2000 of its steps) tiled %dx%dx%d across the periodic cross-section (%d atoms: %d/20 of the "
            "cross-section and 1/4 of the length of the 1 M workload), unmodified reference (LAMMPS 7Feb2024 + "
            "src/OBMD, g++ -O3, MPI stubs, 1 core; no MPI in the image), LAMMPS loop time %.2f s, %d USHER "
            "trials; our arm times the identical workload as `like_for_like`.  The reference's USHER energy is "
            "a brute-force scan of all owned atoms per iteration, so its atom-steps/s FALLS with the tiling: "
            "the 4x5x4 workload itself would take ~%.0f s per step (cost model in bench.py), i.e. ~%.1f k "
            "atom-steps/s" % (steps, reps[0], reps[1], reps[2], n, reps[1] * reps[2], loop, ntrials, full,
                              12008 * 80 / full / 1e3))


def cpu_baseline(budget_s=25.0, steps=4, reps=None):
    if ref_binary() is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/lmp_serial missing (run oracle/build_ref.sh)"}
    reps = reps or pick_sample(steps, budget_s)
    a, loop, ntr = run_reference(reps, steps)
    return {"tiling": list(reps), "value": a / loop, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": sample_text(reps, steps, loop, ntr)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    reps = tuple(args.like_reps) if args.like_reps else pick_sample(total, args.ref_budget)
    if ref_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lmp_serial not built"}))
        return
    # warm-up and timed steps are separate runs of the binary: LAMMPS' loop time excludes setup
    if args.warmup > 0:
        run_reference(reps, args.warmup)
    a, loop, ntr = run_reference(reps, args.steps)
    val = a / loop
    world = args.gpus
    workload = args.workload or ("1m" if world == 1 else "16m")
    if workload == "1m":
        r1 = tuple(args.reps)
        wreps, scaling = (r1[0] * world, r1[1], r1[2]), "weak"
        B = 0.15 * TILE_LX * r1[0]
    else:
        wreps, scaling = TILINGS[workload], "strong"
        B = 0.15 * TILE_LX * wreps[0]
    nt = wreps[0] * wreps[1] * wreps[2]
    kw = dict(regime_kw(args.regime, nt), buffer=B)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * loop / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(wreps, 12008 * nt, args.regime, kw),
                       "sample": sample_text(reps, args.steps, loop, ntr)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                             "sample": sample_text(reps, args.steps, loop, ntr)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reps", type=int, nargs=3, default=[4, 5, 4], help="tiling of the shipped box (1m workload)")
    ap.add_argument("--workload", default=None, choices=["1m", "4m", "16m"],
                    help="default: 1m at --gpus 1 (BASELINE config 2), 16m at --gpus N > 1 (config 5, strong "
                         "scaling); 1m at N > 1 is weak scaling, one tiling per GPU")
    ap.add_argument("--regime", default="R1", choices=["R0", "R1"])
    ap.add_argument("--lammps-steps", type=int, default=500,
                    help="also time the reference LAMMPS binary + plugin (drop-in e2e) for this many steps")
    ap.add_argument("--equil", type=int, default=2000,
                    help="untimed steps of the full step before the timed region (reach steady exchange)")
    ap.add_argument("--like-reps", type=int, nargs=3, default=[1, 1, 2],
                    help="tiling of the reference's R1 snapshot that BOTH arms time (like_for_like / --impl reference)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--secondary-steps", type=int, default=50)
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (r0, strong_16m / weak)")
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
