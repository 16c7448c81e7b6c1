"""Shared helpers for the parity tests: build matching (oracle, device) set-ups of the shipped
OBMD_DPD deck (tests/golden/dpd_8map_obmd.npz) and of synthetic tiled boxes."""
import ctypes as C
import os

import numpy as np

import oracle_api as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")

USHER = (1, 31.03, 1.0, 0.02, 1e4, 1.5, 1.0, 40)


def load(name):
    return np.load(os.path.join(G, name))


def shipped():
    d = load("dpd_8map_obmd.npz")
    return d["tag"].astype(np.int32), d["x"].copy(), d["box"].copy()


def tiled(reps, jitter=0.0, seed=1):
    """Tile the shipped configuration by integer translations (SURVEY 8d synthetic inputs)."""
    tag, x, box = shipped()
    L = box[:, 1] - box[:, 0]
    nx, ny, nz = reps
    parts = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                parts.append(x + np.array([i, j, k]) * L)
    X = np.concatenate(parts)
    if jitter > 0.0:
        X = X + np.random.default_rng(seed).uniform(-jitter, jitter, X.shape)
    B = np.array([[0.0, L[0] * nx], [0.0, L[1] * ny], [0.0, L[2] * nz]])
    X[:, 0] = np.clip(X[:, 0], 0.0, B[0, 1])
    X[:, 1] = np.mod(X[:, 1], B[1, 1])
    X[:, 2] = np.mod(X[:, 2], B[2, 1])
    T = np.arange(1, len(X) + 1, dtype=np.int32)
    return T, X, B


def device_sim(box, *, dt=0.001464, temperature=1.0, rc=1.0, pair_seed=2616, a0=209.6, gamma=4.5,
               masses=(1.0,), fix=True, device=0, **fixkw):
    """ObmdB200 context configured like oracle_api.make_fix(**same kwargs)."""
    import obmd_b200 as ob
    sim = ob.ObmdB200(device)
    sim.set_box(box[:, 0], box[:, 1])
    sim.set_masses(list(masses))
    sim.set_timestep(dt)
    sim.pair_style_dpd(temperature, rc, pair_seed)
    sim.pair_coeff("*", "*", a0, gamma, rc)
    sim.pair_init(1.0)
    if fix:
        sim.fix_obmd(fix_args(box, **fixkw))
    return sim


def fix_args(box, *, ntype=1, nfreq=1, seed=6111, pxx=188.0, pxy=0.0, pxz=0.0, dpxx=0.0, freq=0.0,
             alpha=0.7, tau=0.005, nbuf=1327.0, buffer_size=None, gfac=0.25, maxattempt=1,
             usher=USHER, shear_size=0.0, regions=None):
    import obmd_b200 as ob
    a = ob.FixArgs()
    a.ntype, a.nfreq, a.seed = ntype, nfreq, seed
    a.pxx, a.pxy, a.pxz, a.dpxx, a.freq = pxx, pxy, pxz, dpxx, freq
    a.alpha, a.tau, a.nbuf = alpha, tau, nbuf
    if buffer_size is None:
        buffer_size = 0.15 * box[0, 1]
    a.buffer_size, a.g_fac = buffer_size, gfac
    a.step_parallel, a.step_perp, a.maxattempt = 0, 1, maxattempt
    (a.usherflag, a.etarget, a.ds0, a.dtheta0, a.uovlp, a.dsovlp, a.eps, a.nattempt) = usher
    if regions is None:
        regions = ob.default_regions(box[:, 0], box[:, 1], buffer_size, shear_size)
    for k, (lo, hi) in enumerate(regions):
        for d in range(3):
            a.region[k].lo[d] = lo[d]
            a.region[k].hi[d] = hi[d]
    return a


def oracle_sim(box, x, v=None, tag=None, **kw):
    kw = dict(kw)
    for k in ("fix", "device", "masses"):
        kw.pop(k, None)
    if "usher" not in kw:
        kw["usher"] = USHER
    p = O.make_fix(box[:, 0], box[:, 1], **kw)
    a = O.OracleAtoms(x, v=v, tag=tag)
    return p, a


def mixed_rel_err(f, fr):
    """|df| / max(|f_i|, <|f|>)  -- the tolerance form SURVEY 7.4 prescribes for cancelling sums."""
    scale = np.maximum(np.abs(fr), np.abs(fr).mean() + 1e-300)
    return float(np.max(np.abs(f - fr) / scale))


def smoke_check():
    """Tiny end-to-end check used by __graft_entry__.smoke(): the hooks of one timestep on the
    shipped configuration (12 637 atoms, R1-like parameters so USHER runs) on cuda:0, each
    compared with the oracle fed the same inputs: USHER trials and bookkeeping bit-exact,
    deterministic forces (T = 0) within 1e-9."""
    tag, x, box = shipped()
    rng = np.random.default_rng(7)
    v = rng.normal(0.0, 1.0, x.shape)
    kw = dict(alpha=1.0, nbuf=2000.0, temperature=0.0)
    sim = device_sim(box, **kw)
    sim.upload(x, v, np.ones(len(x), np.int32), tag)
    sim.build_cells()
    rep, tr = sim.pre_exchange(maxtrials=4096)
    sim.pbc()
    sim.pair_compute(1)
    sim.post_force()
    out = sim.download()
    p, a = oracle_sim(box, x, v=v, tag=tag, **kw)
    L = O.lib()
    trials = (O.Trial * 4096)()
    L.orc_fix_pre_exchange(C.byref(p), C.byref(a.c), trials, 4096)
    L.orc_pbc(C.byref(p), C.byref(a.c))
    L.orc_pair_compute(C.byref(p), C.byref(a.c), 0, 1, None, None)
    L.orc_fix_post_force(C.byref(p), C.byref(a.c))
    assert rep.ntrials == p.last_ntrials and rep.naccept == p.last_naccept, (rep.ntrials, p.last_ntrials)
    for k in range(rep.ntrials):
        assert tuple(tr[k].end) == tuple(trials[k].end) and tr[k].accepted == trials[k].accepted, k
    assert np.array_equal(out["tag"], a.view("tag"))
    assert np.array_equal(out["x"], a.view("x"))
    err = mixed_rel_err(out["f"], a.view("f"))
    assert err < 1e-9, err
    sim.close()
