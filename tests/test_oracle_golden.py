"""Pin the CPU oracle (oracle/obmd_oracle.c) against golden vectors produced by the
UNMODIFIED reference binary (tests/golden/make_golden.py).  CPU only."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle_api as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


@pytest.fixture(scope="module")
def shipped():
    d = load("dpd_8map_obmd.npz")
    return d["tag"], d["x"], d["box"]


def sig6(v):
    return float("%.6g" % v)


def test_ranpark_known_answers():
    L = O.lib()
    s = C.c_int(6111)
    seq = [L.orc_ranpark_uniform(C.byref(s)) for _ in range(40)]
    # independent restatement of random_park.cpp:43-50 in exact integer arithmetic
    t = 6111
    for k in range(40):
        t = (16807 * t) % 2147483647
        assert seq[k] == t * (1.0 / 2147483647)
    assert L.orc_ranpark_skip(6111, 40) == t
    assert L.orc_ranpark_skip(6111, 0) == 6111
    assert L.orc_ranpark_skip(12345, 1000003) == (pow(16807, 1000003, 2147483647) * 12345) % 2147483647


def test_ranmars_against_reference_two_atom_deck():
    """fx on atom 1 of a static dimer = one RanMars Gaussian per step (pair_dpd.cpp:122)."""
    L = O.lib()
    fx_ref = load("ranmars_two_atoms.npz")["fx"]
    st = (C.c_char * 2048)()
    L.orc_ranmars_init(st, 2616)
    a0, gamma, rc, dt = 209.6, 4.5, 1.0, 0.001464
    sigma = math.sqrt(2.0 * 1.0 * 1.0 * gamma)
    dtinvsqrt = 1.0 / math.sqrt(dt)
    delx = 5.0 - 5.5
    r = math.sqrt(delx * delx)
    rinv = 1.0 / r
    wd = 1.0 - r / rc
    for n in range(len(fx_ref)):
        xi = L.orc_ranmars_gaussian(st)
        fpair = a0 * wd
        fpair -= gamma * wd * wd * 0.0 * rinv
        fpair *= 1.0
        fpair += 1.0 * sigma * wd * xi * dtinvsqrt
        fpair *= rinv
        assert delx * fpair == fx_ref[n], n


def test_b1_usher_stream_bit_exact(shipped):
    """RanPark trial stream, USHER verdicts, accepted positions and tags."""
    tag, x, box = shipped
    L = O.lib()
    g = load("b1_usher_static.npz")
    for use_cells in (0, 1):
        p = O.make_fix(box[:, 0], box[:, 1], alpha=1.0, nbuf=2000.0, use_cells=use_cells)
        a = O.OracleAtoms(x, tag=tag)
        trials = (O.Trial * 4096)()
        ver = []
        for _ in range(int(g["nsteps"])):
            L.orc_fix_pre_exchange(C.byref(p), C.byref(a.c), trials, 4096)
            for k in range(p.last_ntrials):
                t = trials[k]
                ver.append((t.accepted, sig6(t.energy), t.iter if t.accepted else -1))
        ref = g["verdicts"]
        assert len(ver) == len(ref)
        got = np.array(ver, dtype=np.float64)
        assert np.array_equal(got[:, 0], ref[:, 0])
        assert np.array_equal(got[:, 2], ref[:, 2])
        assert np.allclose(got[:, 1], ref[:, 1], rtol=2e-6, atol=0)
        assert np.array_equal(a.view("tag"), g["final_tag"])
        assert np.array_equal(a.view("x")[len(tag):], g["inserted_x"])  # bit-exact


def test_b9_usher_overlap_branch_bit_exact(shipped):
    """uovlp = 150: the overlap branch of usher() (fix_obmd_merged.cpp:1566-1580, ds = dsovlp -
    pow(4 eps / E, 1/12)) is taken by a large share of the trials; verdicts, accepted positions and
    tags of 10 steps of the unmodified reference binary."""
    tag, x, box = shipped
    L = O.lib()
    g = load("b9_usher_overlap.npz")
    ush = (1, 31.03, 1.0, 0.02, 150.0, 1.5, 1.0, 40)
    for use_cells in (0, 1):
        p = O.make_fix(box[:, 0], box[:, 1], alpha=1.0, nbuf=2000.0, use_cells=use_cells, usher=ush)
        a = O.OracleAtoms(x, tag=tag)
        trials = (O.Trial * 4096)()
        ver = []
        nover = 0
        for _ in range(int(g["nsteps"])):
            L.orc_fix_pre_exchange(C.byref(p), C.byref(a.c), trials, 4096)
            for k in range(p.last_ntrials):
                t = trials[k]
                ver.append((t.accepted, sig6(t.energy), t.iter if t.accepted else -1))
                nover += t.energy > 150.0
        ref = g["verdicts"]
        assert len(ver) == len(ref)
        got = np.array(ver, dtype=np.float64)
        assert np.array_equal(got[:, 0], ref[:, 0])
        assert np.array_equal(got[:, 2], ref[:, 2])
        assert np.allclose(got[:, 1], ref[:, 1], rtol=2e-6, atol=0)
        assert np.array_equal(a.view("tag"), g["final_tag"])
        assert np.array_equal(a.view("x")[len(tag):], g["inserted_x"])  # bit-exact
        assert nover > 5      # trials that END above uovlp; many more pass through the branch on the way


def test_b6_usher_maxattempt3_slot_bookkeeping(shipped):
    """maxattempt 3 (fix_obmd_merged.cpp:912-1172): verdicts WITH the attempt number inside each
    insertion slot, accepted positions and tags, 6 steps of the unmodified reference binary."""
    tag, x, box = shipped
    L = O.lib()
    g = load("b6_usher_maxattempt3.npz")
    p = O.make_fix(box[:, 0], box[:, 1], alpha=1.0, nbuf=2000.0, maxattempt=3)
    a = O.OracleAtoms(x, tag=tag)
    trials = (O.Trial * 8192)()
    ver = []
    for _ in range(int(g["nsteps"])):
        L.orc_fix_pre_exchange(C.byref(p), C.byref(a.c), trials, 8192)
        attempt, side = 0, 0
        for k in range(p.last_ntrials):
            t = trials[k]
            if t.side != side:            # try_inserting is called once per side: slots restart
                attempt, side = 0, t.side
            attempt += 1
            ver.append((t.accepted, sig6(t.energy), attempt))
            if t.accepted or attempt == 3:
                attempt = 0
    ref = g["verdicts"]
    assert len(ver) == len(ref) and ref[:, 2].max() == 3
    got = np.array(ver, dtype=np.float64)
    assert np.array_equal(got[:, 0], ref[:, 0])
    assert np.array_equal(got[:, 2], ref[:, 2])          # attempt number of every verdict line
    assert np.allclose(got[:, 1], ref[:, 1], rtol=2e-6, atol=0)
    assert np.array_equal(a.view("tag"), g["final_tag"])
    assert np.array_equal(a.view("x")[len(tag):], g["inserted_x"])  # bit-exact


def _post_force_once(p, a):
    L = O.lib()
    L.orc_fix_pre_exchange(C.byref(p), C.byref(a.c), None, 0)
    a.f[:] = 0.0
    L.orc_fix_post_force(C.byref(p), C.byref(a.c))


def test_b2_normal_force_bit_exact(shipped):
    tag, x, box = shipped
    g = load("b2_normal_force.npz")
    p = O.make_fix(box[:, 0], box[:, 1], a0=0.0, gamma=0.0, dpxx=10.0, freq=5.0)
    a = O.OracleAtoms(x, tag=tag)
    assert not g["f0"].any()          # setup() forces are exactly zero (Q3)
    for s in (1, 2, 3):
        _post_force_once(p, a)
        assert a.n == len(tag)
        assert np.array_equal(a.view("f"), g["f%d" % s]), s
    ly = box[1, 1] - box[1, 0]; lz = box[2, 1] - box[2, 0]
    left = x[:, 0] < 0.15 * box[0, 1]
    want = (188.0 + 10.0 * math.sin(2 * math.pi * 5.0 * 3 * 0.001464)) * ly * lz
    assert abs(a.view("f")[left, 0].sum() - want) < 1e-8 * want


def test_b4_shear_force_bit_exact(shipped):
    tag, x, box = shipped
    g = load("b4_shear_force.npz")
    p = O.make_fix(box[:, 0], box[:, 1], a0=0.0, gamma=0.0, pxy=2.0, pxz=-1.0, shear_size=2.5)
    a = O.OracleAtoms(x, tag=tag)
    _post_force_once(p, a)
    assert np.array_equal(a.view("f"), g["f1"])


def test_b3_ballistic_bookkeeping_bit_exact(shipped):
    """20 steps of delete / insert / weights / NVE / PBC: order, x, v, f."""
    tag, x, box = shipped
    g = load("b3_ballistic.npz")
    L = O.lib()
    p = O.make_fix(box[:, 0], box[:, 1], a0=0.0, gamma=0.0, dt=0.01, alpha=1.0, tau=0.05,
                   nbuf=2000.0)
    a = O.OracleAtoms(x, v=g["v0"], tag=tag)
    trials = (O.Trial * 4096)()
    ndel = nins = 0
    for s in range(1, 21):
        L.orc_step(C.byref(p), C.byref(a.c), 0, s, trials, 4096)
        ndel += p.last_ndel; nins += p.last_naccept
    assert (ndel, nins) == (int(g["ndel"]), int(g["nins"]))
    assert np.array_equal(a.view("tag"), g["tag20"])
    assert np.array_equal(a.view("x"), g["x20"])
    assert np.array_equal(a.view("v"), g["v20"])
    assert np.array_equal(a.view("f"), g["f20"])


def test_p0_pair_forces_T0(shipped):
    """Deterministic DPD (sigma = 0) forces, energy and pressure tensor vs the reference."""
    tag, x, box = shipped
    g = load("p0_pair_T0.npz")
    L = O.lib()
    p = O.make_fix(box[:, 0], box[:, 1], temperature=0.0)
    a = O.OracleAtoms(x, v=g["v"], tag=tag)
    eng = C.c_double(); vir = (C.c_double * 6)()
    L.orc_pair_compute(C.byref(p), C.byref(a.c), 0, 0, C.byref(eng), vir)
    f = a.view("f"); fr = g["f"]
    scale = np.maximum(np.abs(fr), np.abs(fr).mean())
    assert np.max(np.abs(f - fr) / scale) < 1e-9     # tolerance stated by north_star
    th = g["thermo"]   # step pe press pxx pyy pzz pxy pxz pyz
    assert abs(eng.value - th[1]) < 1e-9 * abs(th[1])
    vol = np.prod(box[:, 1] - box[:, 0])
    v = g["v"]
    ke = [np.sum(v[:, i] * v[:, j]) for i, j in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))]
    pt = [(ke[k] + vir[k]) / vol for k in range(6)]
    assert np.allclose(pt, th[3:9], rtol=1e-9, atol=1e-9)


def test_p1_pair_forces_two_types_T0(shipped):
    """Two atom types with per-pair a0, gamma and cutoff (pair_coeff i j a0 gamma rc,
    pair_dpd.cpp:222-251, init_one :279-291): forces and energy vs the reference binary.  The GPU
    test of the same coefficients (test_gpu_parity.py) compares against this oracle path."""
    tag, x, box = shipped
    g = load("p1_pair_two_types_T0.npz")
    L = O.lib()
    p = O.make_fix(box[:, 0], box[:, 1], ntypes=2, mass=(1.0, 2.0), temperature=0.0)
    for (i, j), (a0, gamma, rc) in {(1, 1): (209.6, 4.5, 1.0), (1, 2): (150.0, 3.0, 0.9),
                                    (2, 2): (100.0, 6.0, 0.8)}.items():
        p.a0[i][j] = a0; p.gamma[i][j] = gamma; p.cut[i][j] = rc
    L.orc_pair_init(C.byref(p))
    a = O.OracleAtoms(x, v=g["v"], tag=tag, type_=g["type"])
    assert 0.3 < (g["type"] == 2).mean() < 0.5
    eng = C.c_double(); vir = (C.c_double * 6)()
    L.orc_pair_compute(C.byref(p), C.byref(a.c), 0, 0, C.byref(eng), vir)
    f = a.view("f"); fr = g["f"]
    scale = np.maximum(np.abs(fr), np.abs(fr).mean())
    assert np.max(np.abs(f - fr) / scale) < 1e-9
    th = g["thermo"]   # step pe press
    assert abs(eng.value - th[1]) < 1e-9 * abs(th[1])


TWO_TYPE_COEFF = {(1, 1): (209.6, 4.5, 1.0), (1, 2): (150.0, 3.0, 0.9), (2, 2): (100.0, 6.0, 0.8)}


def two_type_fix(box, **kw):
    p = O.make_fix(box[:, 0], box[:, 1], ntypes=2, mass=(1.0, 2.0), **kw)
    for (i, j), (a0, gamma, rc) in TWO_TYPE_COEFF.items():
        p.a0[i][j] = a0; p.gamma[i][j] = gamma; p.cut[i][j] = rc
    O.lib().orc_pair_init(C.byref(p))
    return p


def test_b7_usher_two_types_insert_type_2(shipped):
    """USHER with type-dependent pair coefficients: 40 % of the atoms are type 2 and `fix obmd 2 ...`
    inserts type-2 atoms (energy()/single() use a0, cut of the (2, j) pairs, fix_obmd_merged.cpp:1774-1810).
    Verdicts, accepted positions, tags and types vs 8 steps of the unmodified reference binary."""
    tag, x, box = shipped
    g = load("b7_usher_two_types.npz")
    typ0 = load("p1_pair_two_types_T0.npz")["type"]      # same `set type/fraction 2 0.4 12345`
    assert np.array_equal(typ0, g["final_type"][:len(tag)])
    L = O.lib()
    p = two_type_fix(box, alpha=1.0, nbuf=2000.0, ntype=2)
    a = O.OracleAtoms(x, tag=tag, type_=typ0)
    trials = (O.Trial * 4096)()
    ver = []
    for _ in range(int(g["nsteps"])):
        L.orc_fix_pre_exchange(C.byref(p), C.byref(a.c), trials, 4096)
        for k in range(p.last_ntrials):
            t = trials[k]
            ver.append((t.accepted, sig6(t.energy), t.iter if t.accepted else -1))
    ref = g["verdicts"]
    assert len(ver) == len(ref) and ref[:, 0].sum() > 100
    got = np.array(ver, dtype=np.float64)
    assert np.array_equal(got[:, 0], ref[:, 0])
    assert np.array_equal(got[:, 2], ref[:, 2])
    assert np.allclose(got[:, 1], ref[:, 1], rtol=2e-6, atol=0)
    assert np.array_equal(a.view("tag"), g["final_tag"])
    assert np.array_equal(a.view("type"), g["final_type"])
    assert np.array_equal(a.view("x")[len(tag):], g["inserted_x"])  # bit-exact


def test_b8_ballistic_two_masses_bit_exact(shipped):
    """20 ballistic steps with atom types of mass 1 and 2 and tangential flux on shear regions: FixNVE
    with dtf/m, momentum of the deleted atoms, mass-weighted tangential weights, type of the inserted
    atoms -- order, x, v, f bit for bit vs the reference binary."""
    tag, x, box = shipped
    g = load("b8_ballistic_two_masses.npz")
    L = O.lib()
    p = O.make_fix(box[:, 0], box[:, 1], ntypes=2, mass=(1.0, 2.0), a0=0.0, gamma=0.0, dt=0.01, alpha=1.0,
                   tau=0.05, nbuf=2000.0, pxy=2.0, pxz=-1.0, shear_size=2.5)
    a = O.OracleAtoms(x, v=g["v0"], tag=tag, type_=g["type0"])
    trials = (O.Trial * 4096)()
    ndel = nins = 0
    for s in range(1, 21):
        L.orc_step(C.byref(p), C.byref(a.c), 0, s, trials, 4096)
        ndel += p.last_ndel; nins += p.last_naccept
    assert (ndel, nins) == (int(g["ndel"]), int(g["nins"]))
    assert np.array_equal(a.view("tag"), g["tag20"]) and np.array_equal(a.view("type"), g["type20"])
    assert np.array_equal(a.view("x"), g["x20"])
    assert np.array_equal(a.view("v"), g["v20"])
    assert np.array_equal(a.view("f"), g["f20"])


def test_m0_bond_harmonic_vs_reference(shipped):
    """Bonded path (BASELINE config 4, the bonded forces only): orc_bond_harmonic against a run of the
    reference binary in atom_style bond with 6803 harmonic bonds of two types, some across the periodic
    faces (tests/golden/make_golden.py m0): bond-only forces and energy, then bonds + T = 0 DPD forces."""
    tag, x, box = shipped
    g = load("m0_bond_harmonic.npz")
    b = np.ascontiguousarray(g["bonds"], np.int32)
    bt, t1, t2 = (np.ascontiguousarray(b[:, k]) for k in range(3))
    k = np.ascontiguousarray(g["k"]); r0 = np.ascontiguousarray(g["r0"])
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L = O.lib()
    L.orc_bond_harmonic.restype = C.c_double
    p = O.make_fix(box[:, 0], box[:, 1], temperature=0.0)
    a = O.OracleAtoms(x, v=g["v"], tag=tag)
    e = L.orc_bond_harmonic(C.byref(p), C.byref(a.c), len(b), ip(bt), ip(t1), ip(t2), dp(k), dp(r0))
    f = a.view("f")
    assert np.abs(g["f_bond"]).max() > 10.0
    assert np.abs(f - g["f_bond"]).max() <= 1e-9 * np.abs(g["f_bond"]).max()
    assert abs(e - float(g["ebond"])) <= 1e-10 * abs(float(g["ebond"]))
    # bonds + pair forces (special_bonds lj 1 1 1: no exclusions)
    a2 = O.OracleAtoms(x, v=g["v"], tag=tag)
    L.orc_pair_compute(C.byref(p), C.byref(a2.c), 0, 0, None, None)
    L.orc_bond_harmonic(C.byref(p), C.byref(a2.c), len(b), ip(bt), ip(t1), ip(t2), dp(k), dp(r0))
    scale = np.abs(g["f_both"]).max()
    assert np.abs(a2.view("f") - g["f_both"]).max() <= 1e-9 * scale


def test_m1_bond_fene_vs_reference(shipped):
    """orc_bond_fene (MOLECULE/bond_fene.cpp:53-135) against the reference binary on the same 6803 bonds:
    log term and WCA part (bonds shorter than 2^(1/6) sigma exist: |f| up to 2.7e4)."""
    tag, x, box = shipped
    g = load("m1_bond_fene.npz")
    b = np.ascontiguousarray(g["bonds"], np.int32)
    bt, t1, t2 = (np.ascontiguousarray(b[:, k]) for k in range(3))
    co = [np.ascontiguousarray(g[n]) for n in ("k", "r0", "epsilon", "sigma")]
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L = O.lib()
    L.orc_bond_fene.restype = C.c_double
    p = O.make_fix(box[:, 0], box[:, 1], temperature=0.0)
    a = O.OracleAtoms(x, v=g["v"], tag=tag)
    e = L.orc_bond_fene(C.byref(p), C.byref(a.c), len(b), ip(bt), ip(t1), ip(t2), *[dp(c) for c in co])
    scale = np.abs(g["f_bond"]).max()
    assert scale > 1e3
    assert np.abs(a.view("f") - g["f_bond"]).max() <= 1e-9 * scale
    assert abs(e - float(g["ebond"])) <= 1e-10 * abs(float(g["ebond"]))
