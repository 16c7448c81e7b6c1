"""ctypes binding of libobmd_b200.so (include/obmd_b200.h)."""
import ctypes as C
import os

import numpy as np

ABI_VERSION = 1
HERE = os.path.dirname(os.path.abspath(__file__))
IMG0 = (512 << 20) | (512 << 10) | 512


class ObmdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("obmd_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class Block(C.Structure):
    _fields_ = [("lo", C.c_double * 3), ("hi", C.c_double * 3)]


class FixArgs(C.Structure):
    """obmd_fix_args: `fix obmd` arguments after parsing (fix_obmd_merged.cpp:88-164, 1881-2164)."""
    _fields_ = [("ntype", C.c_int), ("nfreq", C.c_int), ("seed", C.c_int),
                ("pxx", C.c_double), ("pxy", C.c_double), ("pxz", C.c_double), ("dpxx", C.c_double),
                ("freq", C.c_double), ("alpha", C.c_double), ("tau", C.c_double), ("nbuf", C.c_double),
                ("region", Block * 6),
                ("buffer_size", C.c_double), ("g_fac", C.c_double),
                ("step_parallel", C.c_int), ("step_perp", C.c_int), ("maxattempt", C.c_int),
                ("usherflag", C.c_int),
                ("etarget", C.c_double), ("ds0", C.c_double), ("dtheta0", C.c_double),
                ("uovlp", C.c_double), ("dsovlp", C.c_double), ("eps", C.c_double),
                ("nattempt", C.c_int)]


class FixState(C.Structure):
    _fields_ = [("rng_seed", C.c_int), ("simulation_time", C.c_double),
                ("momentumForce_left", C.c_double * 3), ("momentumForce_right", C.c_double * 3),
                ("shearForce_left", C.c_double * 3), ("shearForce_right", C.c_double * 3),
                ("ninserted", C.c_longlong), ("ndeleted", C.c_longlong),
                ("ntrials", C.c_longlong), ("nrounds", C.c_longlong)]


class PreExchangeReport(C.Structure):
    _fields_ = [("ndel", C.c_int), ("cnt_left", C.c_int), ("cnt_right", C.c_int),
                ("ninsert_left", C.c_int), ("ninsert_right", C.c_int),
                ("ntrials", C.c_int), ("naccept", C.c_int), ("usher_rounds", C.c_int),
                ("maxtag", C.c_int), ("nlocal", C.c_int),
                ("vnewl", C.c_double * 3), ("vnewr", C.c_double * 3)]


class Trial(C.Structure):
    _fields_ = [("start", C.c_double * 3), ("end", C.c_double * 3), ("energy", C.c_double),
                ("iter", C.c_int), ("accepted", C.c_int), ("side", C.c_int), ("tag", C.c_int)]


def lib_path():
    return os.path.join(HERE, "libobmd_b200.so")


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/obmd_b200.h declares
SIGNATURES = {
    "obmd_abi_version": (C.c_int, []),
    "obmd_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "obmd_destroy": (None, [C.c_void_p]),
    "obmd_last_error": (C.c_char_p, [C.c_void_p]),
    "obmd_set_box": (C.c_int, [C.c_void_p, _dp, _dp]),
    "obmd_set_masses": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "obmd_set_timestep": (C.c_int, [C.c_void_p, C.c_double]),
    "obmd_pair_settings": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int]),
    "obmd_pair_coeff": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "obmd_pair_init": (C.c_int, [C.c_void_p, C.c_double]),
    "obmd_pair_single": (C.c_double, [C.c_void_p, C.c_int, C.c_int, C.c_double, _dp]),
    "obmd_fix_create": (C.c_int, [C.c_void_p, C.POINTER(FixArgs)]),
    "obmd_fix_set_flux": (C.c_int, [C.c_void_p] + [C.c_double] * 8),
    "obmd_fix_get_state": (C.c_int, [C.c_void_p, C.POINTER(FixState)]),
    "obmd_fix_set_state": (C.c_int, [C.c_void_p, C.POINTER(FixState)]),
    "obmd_upload": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _ip, _ip, _ip, C.c_int]),
    "obmd_upload_xv": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "obmd_download": (C.c_int, [C.c_void_p, _ip, _dp, _dp, _dp, _ip, _ip, _ip]),
    "obmd_nlocal": (C.c_int, [C.c_void_p]),
    "obmd_nve_initial": (C.c_int, [C.c_void_p]),
    "obmd_build_cells": (C.c_int, [C.c_void_p]),
    "obmd_pre_exchange": (C.c_int, [C.c_void_p, C.POINTER(PreExchangeReport), C.POINTER(Trial), C.c_int]),
    "obmd_pbc": (C.c_int, [C.c_void_p]),
    "obmd_pair_compute": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int]),
    "obmd_pair_tallies": (C.c_int, [C.c_void_p, _dp, _dp]),
    "obmd_post_force": (C.c_int, [C.c_void_p]),
    "obmd_zero_forces": (C.c_int, [C.c_void_p]),
    "obmd_nve_final": (C.c_int, [C.c_void_p]),
    "obmd_setup": (C.c_int, [C.c_void_p, C.c_longlong]),
    "obmd_run": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int]),
    "obmd_reduce_kinetic": (C.c_int, [C.c_void_p, _dp]),
    "obmd_profile_x": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "obmd_exchange": (C.c_int, [C.c_void_p]),
    "obmd_set_usher_order": (C.c_int, [C.c_void_p, C.c_int]),
    "obmd_profile_x2": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]),
    "obmd_slab_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip]),
    "obmd_bond_style_harmonic": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "obmd_bond_style_fene": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]),
    "obmd_upload_bonds": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip, _ip]),
    "obmd_bond_compute": (C.c_int, [C.c_void_p, C.c_int]),
    "obmd_bond_tallies": (C.c_int, [C.c_void_p, _dp, _dp]),
    "obmd_pair_gaussian_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, _dp]),
    "obmd_comm_unique_id": (C.c_int, [C.c_char_p]),
    "obmd_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "obmd_comm_finalize": (C.c_int, [C.c_void_p]),
    "obmd_comm_slab": (C.c_int, [C.c_void_p, _dp, _dp]),
    "obmd_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "obmd_timing_get": (C.c_int, [C.c_void_p, C.c_char_p, _dp, C.POINTER(C.c_longlong)]),
    "obmd_timing_reset": (C.c_int, [C.c_void_p]),
    "obmd_launch_count": (C.c_longlong, [C.c_void_p]),
    "obmd_stream": (C.c_void_p, [C.c_void_p]),
}


def lib():
    """Load libobmd_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ObmdError(-1, "%s not found: build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (nvcc, sm_100a)" % path)
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.obmd_abi_version() != ABI_VERSION:
        raise ObmdError(-1, "ABI version mismatch")
    _lib = L
    return L


def default_regions(boxlo, boxhi, B, shear_size=0.0):
    """leftB rightB leftshear rightshear insertL insertR built from EXACT box bounds."""
    xl, yl, zl = (float(v) for v in boxlo)
    xh, yh, zh = (float(v) for v in boxhi)
    leftB = ((xl, yl, zl), (xl + B, yh, zh))
    rightB = ((xh - B, yl, zl), (xh, yh, zh))
    if shear_size > 0.0:
        lsh = ((xl, yl, zl), (xl + shear_size, yh, zh))
        rsh = ((xh - shear_size, yl, zl), (xh, yh, zh))
    else:
        lsh = rsh = ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    return [leftB, rightB, lsh, rsh, leftB, rightB]


def slab_columns(ncols, nranks, rank, edge_cols=0):
    """Owned cell columns [c0, c1) of `rank`: obmd_slab_partition, the function setup_grid() (csrc/api.cu)
    itself calls (host arithmetic inside the shipped library: loads and runs without a GPU)."""
    c0, c1 = C.c_int(0), C.c_int(0)
    rc = lib().obmd_slab_partition(ncols, nranks, rank, edge_cols, C.byref(c0), C.byref(c1))
    if rc != 0:
        raise ObmdError(rc, "obmd_slab_partition(%d, %d, %d, %d)" % (ncols, nranks, rank, edge_cols))
    return c0.value, c1.value


def slab_bounds(boxlo_x, boxhi_x, cutmax, nranks, rank):
    """x-range [lo, hi) of the slab of `rank` (cell size = Lx / floor(Lx / cutmax))."""
    L = boxhi_x - boxlo_x
    ncols = max(int(L // cutmax), 1)
    c0, c1 = slab_columns(ncols, nranks, rank)
    inv = ncols / L
    return boxlo_x + c0 / inv, (boxhi_x if c1 >= ncols else boxlo_x + c1 / inv)


def slab_mask(x, lo, hi, is_last, boxhi_x):
    """Atoms a rank uploads: x in [lo, hi); the last rank also takes x == boxhi."""
    m = (x >= lo) & (x < hi)
    if is_last:
        m |= (x >= lo) & (x <= boxhi_x)
    return m


def parse_fix_obmd(argv, regions, boxlo, boxhi):
    """Mirror of FixObmdMerged's argument parsing (fix_obmd_merged.cpp:66-164 and options()
    :1881-2164) for ATOM mode: argv is the token list after `fix ID group obmd`, `regions` maps
    region IDs to ((xlo,ylo,zlo),(xhi,yhi,zhi)).  Same keywords, defaults and error strings."""
    if len(argv) < 11:
        raise ObmdError(-2, "Illegal fix obmd command (check # of arguments)")
    a = FixArgs()
    a.ntype, a.nfreq, a.seed = int(argv[0]), int(argv[1]), int(argv[2])
    if a.seed <= 0:
        raise ObmdError(-2, "Illegal fix obmd command")
    vals = []
    for tok in argv[3:11]:
        if tok.startswith("v_"):
            raise ObmdError(-2, "variable-style fix obmd arguments are evaluated by the host "
                            "(use ObmdB200.set_flux each step)")
        vals.append(float(tok))
    a.pxx, a.pxy, a.pxz, a.dpxx, a.freq, a.alpha, a.tau, a.nbuf = vals
    # defaults, :1883-1939
    a.g_fac = 0.25
    a.buffer_size = 0.30 * (boxhi[0] - boxlo[0])
    a.etarget, a.ds0, a.dtheta0, a.uovlp, a.dsovlp, a.eps, a.nattempt = 3.6, 0.1, 0.35, 10000.0, 3.0, 0.15, 40
    a.maxattempt = 1
    a.step_parallel, a.step_perp = 0, 1
    a.usherflag = 0
    nearflag = 0
    ids = [None] * 6
    rest = list(argv[11:])
    i = 0
    while i < len(rest):
        k = rest[i]
        if k in ("region1", "region2", "region3", "region4", "region5", "region6"):
            rid = rest[i + 1]
            if rid not in regions:
                raise ObmdError(-2, "Region ID %s for fix obmd does not exist" % rid)
            ids[int(k[-1]) - 1] = rid
            i += 2
        elif k == "buffersize":
            a.buffer_size = float(rest[i + 1])
            if a.buffer_size <= 0.0:
                raise ObmdError(-2, "Illegal fix obmd command. Parameter buffersize should be larger than 0.0")
            i += 2
        elif k == "gfac":
            a.g_fac = float(rest[i + 1])
            if a.g_fac < 0.0 or a.g_fac > 1.0:
                raise ObmdError(-2, "Illegal fix obmd command. Parameter gfac should be between 0.0 and 1.0")
            i += 2
        elif k == "stepparallel":
            a.step_parallel = int(rest[i + 1])
            if a.step_parallel != 0:
                raise ObmdError(-2, "For now, only the smooth transition distribution function is implemented.")
            i += 2
        elif k == "stepperp":
            a.step_perp = int(rest[i + 1])
            if a.step_perp != 1:
                raise ObmdError(-2, "For now, only the heaviside step distribution function is implemented.")
            i += 2
        elif k == "maxattempt":
            a.maxattempt = int(rest[i + 1])
            i += 2
        elif k == "usher":
            if i + 9 > len(rest):
                raise ObmdError(-2, "Illegal fix obmd command")
            a.usherflag = int(rest[i + 1])
            if a.usherflag == 1 and nearflag == 1:
                raise ObmdError(-2, "You can not have both usher and near")
            a.etarget, a.ds0, a.dtheta0, a.uovlp, a.dsovlp, a.eps = (float(t) for t in rest[i + 2:i + 8])
            a.nattempt = int(rest[i + 8])
            i += 9
        elif k == "charged":
            if int(rest[i + 1]) == 1:
                raise ObmdError(-2, "You can not use charged without MOLECULE mode")
            i += 2
        elif k == "near":
            nearflag = int(rest[i + 1])
            raise ObmdError(-2, "fix obmd near: only usher insertion is implemented on the device path")
        elif k in ("mol", "molfrac", "rigid", "shake", "id", "global", "local", "rate", "vx", "vy",
                   "vz", "orient", "units", "gaussian", "target"):
            raise ObmdError(-2, "fix obmd keyword '%s' is outside the ATOM/usher hot path "
                            "(DESIGN.md: out of scope)" % k)
        else:
            raise ObmdError(-2, "Illegal fix obmd command")
    if a.usherflag == 0 and nearflag == 0:
        raise ObmdError(-2, "You must pick either usher or near")
    if ids[0] is None or ids[1] is None:
        raise ObmdError(-2, "Must specify a region in fix obmd")
    for r in range(6):
        if ids[r] is None:
            raise ObmdError(-2, "Region ID for fix obmd does not exist")
        lo, hi = regions[ids[r]]
        for d in range(3):
            a.region[r].lo[d] = lo[d]
            a.region[r].hi[d] = hi[d]
    return a


def _dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class ObmdB200:
    """One device context = one LAMMPS rank's fix obmd + pair dpd (+ fix nve) state."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.obmd_create(C.byref(h), device)
        if rc != 0:
            raise ObmdError(rc, self.L.obmd_last_error(None).decode())
        self.h = h
        self.ntypes = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.obmd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise ObmdError(rc, self.L.obmd_last_error(self.h).decode())

    # ---- domain / styles -------------------------------------------------------------
    def set_box(self, boxlo, boxhi):
        lo = (C.c_double * 3)(*[float(v) for v in boxlo])
        hi = (C.c_double * 3)(*[float(v) for v in boxhi])
        self._ck(self.L.obmd_set_box(self.h, lo, hi))
        self.boxlo, self.boxhi = tuple(lo), tuple(hi)

    def set_masses(self, masses):
        m = (C.c_double * (len(masses) + 1))(0.0, *[float(v) for v in masses])
        self._ck(self.L.obmd_set_masses(self.h, len(masses), m))
        self.ntypes = len(masses)

    def set_timestep(self, dt):
        self._ck(self.L.obmd_set_timestep(self.h, dt))

    def pair_style_dpd(self, temperature, cut_global, seed):
        """pair_style dpd T rc seed"""
        self._ck(self.L.obmd_pair_settings(self.h, temperature, cut_global, seed))

    def pair_coeff(self, itype, jtype, a0, gamma, cut=0.0):
        """pair_coeff i j a0 gamma [rc]; itype/jtype may be '*'"""
        ir = range(1, self.ntypes + 1) if itype == "*" else [int(itype)]
        jr = range(1, self.ntypes + 1) if jtype == "*" else [int(jtype)]
        for i in ir:
            for j in jr:
                if j >= i:
                    self._ck(self.L.obmd_pair_coeff(self.h, i, j, a0, gamma, cut))

    def pair_init(self, boltz=1.0):
        self._ck(self.L.obmd_pair_init(self.h, boltz))

    def fix_obmd(self, args):
        self._ck(self.L.obmd_fix_create(self.h, C.byref(args)))

    def set_flux(self, pxx, pxy, pxz, dpxx, freq, alpha, tau, nbuf):
        self._ck(self.L.obmd_fix_set_flux(self.h, pxx, pxy, pxz, dpxx, freq, alpha, tau, nbuf))

    def fix_state(self):
        s = FixState()
        self._ck(self.L.obmd_fix_get_state(self.h, C.byref(s)))
        return s

    def set_fix_state(self, s):
        self._ck(self.L.obmd_fix_set_state(self.h, C.byref(s)))

    # ---- atoms -----------------------------------------------------------------------
    def upload(self, x, v, type_, tag, image=None, f=None, capacity=0):
        x = np.ascontiguousarray(x, np.float64); v = np.ascontiguousarray(v, np.float64)
        type_ = np.ascontiguousarray(type_, np.int32); tag = np.ascontiguousarray(tag, np.int32)
        image = None if image is None else np.ascontiguousarray(image, np.int32)
        f = None if f is None else np.ascontiguousarray(f, np.float64)
        self._ck(self.L.obmd_upload(self.h, len(x), _dptr(x), _dptr(v), _dptr(f), _iptr(type_),
                                    _iptr(tag), _iptr(image), int(capacity)))

    def upload_xv(self, x, v):
        x = np.ascontiguousarray(x, np.float64); v = np.ascontiguousarray(v, np.float64)
        self._ck(self.L.obmd_upload_xv(self.h, len(x), _dptr(x), _dptr(v)))

    def download(self, want=("x", "v", "f", "type", "tag", "image"), out=None):
        """Atom arrays in LAMMPS local order.  `out` may hold preallocated (e.g. pinned) arrays with
        room for nlocal rows under the same keys; views of their first nlocal rows are returned."""
        n = self.nlocal
        arr = {}
        for k in want:
            if out is not None and k in out:
                a = out[k][:n]
                assert a.flags.c_contiguous and a.dtype == (np.float64 if k in ("x", "v", "f") else np.int32)
                arr[k] = a
            else:
                arr[k] = np.empty((n, 3)) if k in ("x", "v", "f") else np.empty(n, np.int32)
        out = {}
        nn = C.c_int()
        self._ck(self.L.obmd_download(self.h, C.byref(nn), _dptr(arr.get("x")), _dptr(arr.get("v")),
                                      _dptr(arr.get("f")), _iptr(arr.get("type")), _iptr(arr.get("tag")),
                                      _iptr(arr.get("image"))))
        assert nn.value == n
        out.update(arr)
        return out

    @property
    def nlocal(self):
        return self.L.obmd_nlocal(self.h)

    # ---- hooks -----------------------------------------------------------------------
    # ---- bonded path (bond_style harmonic; static topology) -----------------------------
    def bond_style_harmonic(self, k, r0):
        """k, r0: per bond type, 1-based lists WITHOUT the unused entry 0 (bond_coeff t K r0)."""
        kk = np.ascontiguousarray([0.0] + list(k), np.float64)
        rr = np.ascontiguousarray([0.0] + list(r0), np.float64)
        self._ck(self.L.obmd_bond_style_harmonic(self.h, len(kk) - 1, _dptr(kk), _dptr(rr)))

    def bond_style_fene(self, k, r0, epsilon, sigma):
        """per bond type, lists WITHOUT the unused entry 0 (bond_coeff t K R0 epsilon sigma)."""
        arr = [np.ascontiguousarray([0.0] + list(v), np.float64) for v in (k, r0, epsilon, sigma)]
        self._ck(self.L.obmd_bond_style_fene(self.h, len(arr[0]) - 1, *[_dptr(a) for a in arr]))

    def upload_bonds(self, btype, tag1, tag2):
        t = np.ascontiguousarray(btype, np.int32); a = np.ascontiguousarray(tag1, np.int32)
        b = np.ascontiguousarray(tag2, np.int32)
        self._ck(self.L.obmd_upload_bonds(self.h, len(t), _iptr(t), _iptr(a), _iptr(b)))

    def bond_compute(self, eflag=0):
        self._ck(self.L.obmd_bond_compute(self.h, eflag))

    def bond_tallies(self):
        """(bond energy, virial[6]) of the latest bond_compute(eflag = 1)."""
        e = C.c_double(0.0)
        v = (C.c_double * 6)()
        self._ck(self.L.obmd_bond_tallies(self.h, C.byref(e), v))
        return e.value, np.array(list(v))

    def bond_energy(self):
        return self.bond_tallies()[0]

    def zero_forces(self): self._ck(self.L.obmd_zero_forces(self.h))
    def nve_initial(self): self._ck(self.L.obmd_nve_initial(self.h))
    def nve_final(self): self._ck(self.L.obmd_nve_final(self.h))
    def build_cells(self): self._ck(self.L.obmd_build_cells(self.h))
    def pbc(self): self._ck(self.L.obmd_pbc(self.h))
    def post_force(self): self._ck(self.L.obmd_post_force(self.h))

    def pre_exchange(self, maxtrials=0):
        rep = PreExchangeReport()
        tr = (Trial * maxtrials)() if maxtrials > 0 else None
        self._ck(self.L.obmd_pre_exchange(self.h, C.byref(rep), tr, maxtrials))
        if maxtrials > 0:
            return rep, [tr[k] for k in range(min(rep.ntrials, maxtrials))]
        return rep, []

    def pair_compute(self, ntimestep=0, eflag=0, vflag=0):
        self._ck(self.L.obmd_pair_compute(self.h, ntimestep, eflag, vflag))

    def pair_tallies(self):
        e = C.c_double(); v = (C.c_double * 6)()
        self._ck(self.L.obmd_pair_tallies(self.h, C.byref(e), v))
        return e.value, np.array(list(v))

    def step(self, ntimestep, maxtrials=0):
        """One timestep through the separate hooks, in Verlet::run order."""
        self.nve_initial()
        self.build_cells()
        rep, tr = self.pre_exchange(maxtrials)
        self.pbc()
        self.pair_compute(ntimestep)
        self.post_force()
        self.nve_final()
        return rep, tr

    def setup(self, ntimestep=0):
        """Verlet::setup (verlet.cpp:92-163): [exchange, borders,] neighbour build, pair forces,
        fix setup -> post_force with the (zero-initialised, Q3) flux forces."""
        self._ck(self.L.obmd_setup(self.h, ntimestep))

    # ---- slab decomposition ----------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        rc = lib().obmd_comm_unique_id(buf)
        if rc != 0:
            raise ObmdError(rc, lib().obmd_last_error(None).decode())
        return buf.raw

    def comm_init(self, rank, nranks, uid):
        self._ck(self.L.obmd_comm_init(self.h, rank, nranks, uid))

    def comm_finalize(self):
        self._ck(self.L.obmd_comm_finalize(self.h))

    def comm_slab(self):
        lo = C.c_double(); hi = C.c_double()
        self._ck(self.L.obmd_comm_slab(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def exchange(self): self._ck(self.L.obmd_exchange(self.h))
    def set_usher_order(self, by_tag): self._ck(self.L.obmd_set_usher_order(self.h, int(by_tag)))

    def run(self, first_step, nsteps):
        self._ck(self.L.obmd_run(self.h, first_step, nsteps))

    def kinetic(self):
        out = (C.c_double * 5)()
        self._ck(self.L.obmd_reduce_kinetic(self.h, out))
        return np.array(list(out))

    def profile_x(self, nbins):
        cnt = np.zeros(nbins); mvx = np.zeros(nbins)
        self._ck(self.L.obmd_profile_x(self.h, nbins, _dptr(cnt), _dptr(mvx)))
        return cnt, mvx

    def profile_x2(self, nbins, stress=True):
        """count, sum m vx, sum m v^2 [, sum of per-atom stress xx] per x-bin (obmd_profile_x2)."""
        cnt = np.zeros(nbins); mvx = np.zeros(nbins); mv2 = np.zeros(nbins); sxx = np.zeros(nbins)
        self._ck(self.L.obmd_profile_x2(self.h, nbins, _dptr(cnt), _dptr(mvx), _dptr(mv2),
                                        _dptr(sxx) if stress else None))
        return cnt, mvx, mv2, sxx

    def gaussian_sample(self, n, tag0=1, step=1):
        out = np.zeros(n)
        self._ck(self.L.obmd_pair_gaussian_sample(self.h, n, tag0, step, _dptr(out)))
        return out

    # ---- timing ----------------------------------------------------------------------
    def timing(self, on=True): self._ck(self.L.obmd_timing_enable(self.h, int(on)))
    def timing_reset(self): self._ck(self.L.obmd_timing_reset(self.h))

    def timing_get(self, name):
        ms = C.c_double(); n = C.c_longlong()
        self._ck(self.L.obmd_timing_get(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launches(self):
        return self.L.obmd_launch_count(self.h)
