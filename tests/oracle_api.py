"""ctypes view of oracle/liborc.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
MAXT = 8
IMG0 = (512 << 20) | (512 << 10) | 512

T2 = (C.c_double * (MAXT + 1)) * (MAXT + 1)


class Block(C.Structure):
    _fields_ = [("lo", C.c_double * 3), ("hi", C.c_double * 3)]


class Atoms(C.Structure):
    _fields_ = [("nlocal", C.c_int), ("nmax", C.c_int), ("natoms", C.c_long),
                ("x", C.POINTER(C.c_double)), ("v", C.POINTER(C.c_double)),
                ("f", C.POINTER(C.c_double)),
                ("type", C.POINTER(C.c_int)), ("tag", C.POINTER(C.c_int)),
                ("image", C.POINTER(C.c_int))]


class Fix(C.Structure):
    _fields_ = [
        ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3),
        ("ntypes", C.c_int), ("mass", C.c_double * (MAXT + 1)),
        ("temperature", C.c_double), ("cut_global", C.c_double), ("boltz", C.c_double),
        ("pair_seed", C.c_int),
        ("a0", T2), ("gamma", T2), ("sigma", T2), ("cut", T2), ("cutsq", T2),
        ("dt", C.c_double),
        ("ntype", C.c_int), ("nfreq", C.c_int), ("seed", C.c_int),
        ("pxx", C.c_double), ("pxy", C.c_double), ("pxz", C.c_double), ("dpxx", C.c_double),
        ("freq", C.c_double), ("alpha", C.c_double), ("tau", C.c_double), ("nbuf", C.c_double),
        ("region", Block * 6),
        ("buffer_size", C.c_double), ("g_fac", C.c_double),
        ("maxattempt", C.c_int), ("usherflag", C.c_int),
        ("etarget", C.c_double), ("ds0", C.c_double), ("dtheta0", C.c_double),
        ("uovlp", C.c_double), ("dsovlp", C.c_double), ("eps", C.c_double),
        ("nattempt", C.c_int),
        ("rng_seed", C.c_int), ("simulation_time", C.c_double),
        ("momentumForce_left", C.c_double * 3), ("momentumForce_right", C.c_double * 3),
        ("shearForce_left", C.c_double * 3), ("shearForce_right", C.c_double * 3),
        ("vnewl", C.c_double * 3), ("vnewr", C.c_double * 3),
        ("ninserted", C.c_long), ("ndeleted", C.c_long), ("maxtag_all", C.c_int),
        ("last_cnt_left", C.c_int), ("last_cnt_right", C.c_int),
        ("last_ninsert_left", C.c_int), ("last_ninsert_right", C.c_int),
        ("last_ndel", C.c_int), ("last_ntrials", C.c_int), ("last_naccept", C.c_int),
        ("use_cells", C.c_int),
    ]


class Trial(C.Structure):
    _fields_ = [("start", C.c_double * 3), ("end", C.c_double * 3), ("energy", C.c_double),
                ("iter", C.c_int), ("accepted", C.c_int), ("side", C.c_int), ("tag", C.c_int)]


_lib = None


def lib():
    """Load (building if needed) oracle/liborc.so."""
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ORC_DIR, "liborc.so")
    src = os.path.join(ORC_DIR, "obmd_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORC_DIR, "liborc.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    L.orc_ranpark_uniform.restype = C.c_double
    L.orc_ranpark_uniform.argtypes = [C.POINTER(C.c_int)]
    L.orc_ranpark_skip.restype = C.c_int
    L.orc_ranpark_skip.argtypes = [C.c_int, C.c_long]
    L.orc_ranmars_init.argtypes = [C.c_void_p, C.c_int]
    L.orc_ranmars_uniform.restype = C.c_double
    L.orc_ranmars_uniform.argtypes = [C.c_void_p]
    L.orc_ranmars_gaussian.restype = C.c_double
    L.orc_ranmars_gaussian.argtypes = [C.c_void_p]
    L.orc_pair_init.argtypes = [C.POINTER(Fix)]
    L.orc_pair_single.restype = C.c_double
    L.orc_pair_single.argtypes = [C.POINTER(Fix), C.c_int, C.c_int, C.c_double,
                                  C.POINTER(C.c_double)]
    L.orc_pair_gaussian.restype = C.c_double
    L.orc_pair_gaussian.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint]
    L.orc_pair_compute.argtypes = [C.POINTER(Fix), C.POINTER(Atoms), C.c_int, C.c_uint,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_fix_init.argtypes = [C.POINTER(Fix)]
    L.orc_fix_pre_exchange.argtypes = [C.POINTER(Fix), C.POINTER(Atoms), C.POINTER(Trial), C.c_int]
    L.orc_fix_post_force.argtypes = [C.POINTER(Fix), C.POINTER(Atoms)]
    L.orc_g_par_local.restype = C.c_double
    L.orc_g_par_local.argtypes = [C.POINTER(Fix), C.c_double, C.c_double]
    L.orc_usher.restype = C.c_double
    L.orc_usher.argtypes = [C.POINTER(Fix), C.POINTER(Atoms), C.POINTER(Block),
                            C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.orc_nve_initial.argtypes = [C.POINTER(Fix), C.POINTER(Atoms)]
    L.orc_nve_final.argtypes = [C.POINTER(Fix), C.POINTER(Atoms)]
    L.orc_pbc.argtypes = [C.POINTER(Fix), C.POINTER(Atoms)]
    L.orc_step.argtypes = [C.POINTER(Fix), C.POINTER(Atoms), C.c_int, C.c_uint,
                           C.POINTER(Trial), C.c_int]
    _lib = L
    return L


class OracleAtoms:
    """numpy-backed LAMMPS-layout atom arrays handed to the oracle."""

    def __init__(self, x, v=None, tag=None, type_=None, image=None, f=None, extra=0.25):
        n = len(x)
        nmax = int(n * (1.0 + extra)) + 1024
        self.nmax = nmax
        self.x = np.zeros((nmax, 3)); self.x[:n] = x
        self.v = np.zeros((nmax, 3))
        if v is not None: self.v[:n] = v
        self.f = np.zeros((nmax, 3))
        if f is not None: self.f[:n] = f
        self.tag = np.zeros(nmax, np.int32)
        self.tag[:n] = np.arange(1, n + 1) if tag is None else tag
        self.type = np.zeros(nmax, np.int32)
        self.type[:n] = 1 if type_ is None else type_
        self.image = np.zeros(nmax, np.int32)
        self.image[:n] = IMG0 if image is None else image
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self.c = Atoms(n, nmax, n, dp(self.x), dp(self.v), dp(self.f),
                       ip(self.type), ip(self.tag), ip(self.image))

    @property
    def n(self):
        return self.c.nlocal

    def view(self, name):
        return getattr(self, name)[: self.n]


def make_fix(boxlo, boxhi, *, dt=0.001464, ntypes=1, mass=(1.0,), temperature=1.0, rc=1.0,
             pair_seed=2616, a0=209.6, gamma=4.5, ntype=1, nfreq=1, seed=6111,
             pxx=188.0, pxy=0.0, pxz=0.0, dpxx=0.0, freq=0.0, alpha=0.7, tau=0.005, nbuf=1327.0,
             buffer_size=None, gfac=0.25, maxattempt=1, usher=(1, 31.03, 1.0, 0.02, 1e4, 1.5, 1.0, 40),
             regions=None, shear_size=0.0, use_cells=0):
    """Parameter block equivalent to the shipped deck with regions built from exact box bounds."""
    L = lib()
    p = Fix()
    for d in range(3):
        p.boxlo[d] = boxlo[d]; p.boxhi[d] = boxhi[d]
    p.ntypes = ntypes
    for t in range(ntypes):
        p.mass[t + 1] = mass[t] if t < len(mass) else mass[-1]
    p.temperature = temperature; p.cut_global = rc; p.boltz = 1.0; p.pair_seed = pair_seed
    for i in range(1, ntypes + 1):
        for j in range(i, ntypes + 1):
            p.a0[i][j] = a0; p.gamma[i][j] = gamma; p.cut[i][j] = rc
    p.dt = dt
    L.orc_pair_init(C.byref(p))
    p.ntype = ntype; p.nfreq = nfreq; p.seed = seed
    p.pxx = pxx; p.pxy = pxy; p.pxz = pxz; p.dpxx = dpxx; p.freq = freq
    p.alpha = alpha; p.tau = tau; p.nbuf = nbuf
    if buffer_size is None:
        buffer_size = 0.15 * boxhi[0]
    p.buffer_size = buffer_size; p.g_fac = gfac; p.maxattempt = maxattempt
    (p.usherflag, p.etarget, p.ds0, p.dtheta0, p.uovlp, p.dsovlp, p.eps, p.nattempt) = usher
    if regions is None:
        regions = default_regions(boxlo, boxhi, buffer_size, shear_size)
    for k, (lo, hi) in enumerate(regions):
        for d in range(3):
            p.region[k].lo[d] = lo[d]; p.region[k].hi[d] = hi[d]
    p.use_cells = use_cells
    L.orc_fix_init(C.byref(p))
    return p


def default_regions(boxlo, boxhi, B, shear_size=0.0):
    """leftB rightB leftshear rightshear insertL insertR from EXACT box bounds.
    Arithmetic mirrors what the parity decks hand to LAMMPS: xhi - B evaluated in double."""
    xl, yl, zl = boxlo
    xh, yh, zh = boxhi
    leftB = ((xl, yl, zl), (xl + B, yh, zh))
    rightB = ((xh - B, yl, zl), (xh, yh, zh))
    if shear_size > 0.0:
        lsh = ((xl, yl, zl), (xl + shear_size, yh, zh))
        rsh = ((xh - shear_size, yl, zl), (xh, yh, zh))
    else:
        lsh = ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
        rsh = ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    return [leftB, rightB, lsh, rsh, leftB, rightB]
