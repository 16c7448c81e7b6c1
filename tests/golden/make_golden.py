#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference binary (oracle/_ref/lmp_serial, built by oracle/build_ref.sh from
/root/reference) on small parity decks.  Run in the build container only:

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The decks follow SURVEY.md Appendix B (B1 USHER stream, B2/B5 normal force,
B4 tangential force, B3 ballistic bookkeeping) plus a T=0 DPD force deck and a
two-atom RanMars deck.  All buffer regions are built from the EXACT box bounds
(SURVEY Q13), written as %.17g literals computed here (LAMMPS' own variable
evaluation of 0.15*xhi differs from IEEE by 1 ulp) and atom sorting is disabled (`atom_modify sort 0 0`) so dump row
order == LAMMPS local index order.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LMP = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
DATA = "/root/reference/examples/OBMD_DPD/dpd_8map_obmd.data"

NOISE = re.compile(r"^(FixNVE|linear_momentum|nlocal:|v\[i\]|f\[i\]|x\[i\]|atom->tag|dtf|sumforces|"
                   r"pre_exchange|g_perp_global|avec :::|modify :::|Atom::map_init)")

HEAD = """
units           lj
boundary        f p p
atom_style      atomic
atom_modify     sort 0 0
comm_modify     vel yes
newton          on
read_data       {data}
region          leftB block {xlo:.17g} {xlB:.17g} {ylo:.17g} {yhi:.17g} {zlo:.17g} {zhi:.17g}
region          rightB block {xrB:.17g} {xhi:.17g} {ylo:.17g} {yhi:.17g} {zlo:.17g} {zhi:.17g}
region          leftshear block {ls}
region          rightshear block {rs}
pair_style      dpd {T} 1.0 2616
pair_coeff      * * {a0} {gamma} 1.0
neighbor        0.4 bin
neigh_modify    delay 0 every 1
timestep        {dt}
"""

# Box of the shipped data file and the buffer length, as Python doubles.  Every
# bound is written into the decks with %.17g so that LAMMPS parses exactly the
# doubles the tests hand to the oracle / the CUDA path (tests/oracle_api.py
# default_regions does the same arithmetic).
BOX = ((0.0, 33.59462486002239), (0.0, 11.198208286674133), (0.0, 11.198208286674133))
BUF = 0.15 * BOX[0][1]

FIX = ("fix 2 all obmd 1 1 6111 {pxx} {pxy} {pxz} {dpxx} {freq} {alpha} {tau} {nbuf} "
       "region1 leftB region2 rightB region3 leftshear region4 rightshear "
       "region5 leftB region6 rightB buffersize {B:.17g} gfac 0.25 stepparallel 0 stepperp 1 "
       "maxattempt {maxattempt} usher 1 31.03 1.0 0.02 10000 1.5 1.0 40 charged 0\n")


def run_deck(name, text, workdir):
    path = os.path.join(workdir, "in." + name)
    with open(path, "w") as fh:
        fh.write(text)
    out = os.path.join(workdir, "out." + name)
    with open(out, "w") as fh:
        subprocess.check_call([LMP, "-in", path, "-log", "none"], cwd=workdir, stdout=fh,
                              stderr=subprocess.STDOUT)
    with open(out) as fh:
        return [ln.rstrip("\n") for ln in fh if not NOISE.match(ln)]


def read_dump(path):
    """-> list of (step, ndarray rows) for a `dump custom` text file."""
    frames = []
    with open(path) as fh:
        lines = fh.read().split("\n")
    i = 0
    while i < len(lines):
        if lines[i].startswith("ITEM: TIMESTEP"):
            step = int(lines[i + 1]); n = int(lines[i + 3])
            assert lines[i + 8].startswith("ITEM: ATOMS")
            rows = np.array([ln.split() for ln in lines[i + 9:i + 9 + n]], dtype=np.float64)
            frames.append((step, rows))
            i += 9 + n
        else:
            i += 1
    return frames


def head(**kw):
    (xlo, xhi), (ylo, yhi), (zlo, zhi) = BOX
    S = kw.pop("shear", 0.0)
    d = dict(data=DATA, T=1.0, a0=209.6, gamma=4.5, dt=0.001464, xlo=xlo, xhi=xhi, ylo=ylo,
             yhi=yhi, zlo=zlo, zhi=zhi, xlB=xlo + BUF, xrB=xhi - BUF)
    if S > 0.0:
        d["ls"] = "%.17g %.17g %.17g %.17g %.17g %.17g" % (xlo, xlo + S, ylo, yhi, zlo, zhi)
        d["rs"] = "%.17g %.17g %.17g %.17g %.17g %.17g" % (xhi - S, xhi, ylo, yhi, zlo, zhi)
    else:
        d["ls"] = d["rs"] = "0.0 0.0 0.0 0.0 0.0 0.0"
    d.update(kw)
    return HEAD.format(**d)


def fix(**kw):
    d = dict(pxx=188.0, pxy=0, pxz=0, dpxx=0, freq=0, alpha=0.7, tau=0.005, nbuf=1327, maxattempt=1,
             B=BUF)
    d.update(kw)
    return FIX.format(**d)


def parse_usher(lines):
    """USHER verdict lines -> (accepted, E as printed, iterations or -1)."""
    out = []
    for ln in lines:
        m = re.match(r"USHER accepts at E = (\S+) in attempt No\. \d+ with (\d+) iterations", ln)
        if m:
            out.append((1, float(m.group(1)), int(m.group(2)))); continue
        m = re.match(r"USHER denies at E = (\S+) at attempt", ln)
        if m:
            out.append((0, float(m.group(1)), -1))
    return np.array(out, dtype=np.float64).reshape(-1, 3)


def parse_attempts(lines):
    """USHER verdict lines -> (accepted, E as printed, attempt number within its insertion slot)."""
    out = []
    for ln in lines:
        m = re.match(r"USHER accepts at E = (\S+) in attempt No\. (\d+) with \d+ iterations", ln)
        if m:
            out.append((1, float(m.group(1)), int(m.group(2)))); continue
        m = re.match(r"USHER denies at E = (\S+) at attempt No\. (\d+)", ln)
        if m:
            out.append((0, float(m.group(1)), int(m.group(2))))
    return np.array(out, dtype=np.float64).reshape(-1, 3)


def make_b6(wd, n0):
    """B6: the B1 deck with maxattempt 3 -- pins the slot/attempt bookkeeping of try_inserting
    (fix_obmd_merged.cpp:912-1172): which draw of the RanPark stream belongs to which slot."""
    deck = (head() + fix(alpha=1.0, nbuf=2000, maxattempt=3) +
            "dump d all custom 6 dump.b6 id x y z\ndump_modify d format float %.17g\n"
            "thermo 1\nthermo_style custom step atoms\nrun 6\n")
    lines = run_deck("b6", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b6"))
    step, rows = frames[-1]
    assert step == 6
    ver = parse_attempts(lines)
    np.savez_compressed(os.path.join(HERE, "b6_usher_maxattempt3.npz"),
                        verdicts=ver, final_tag=rows[:, 0].astype(np.int32),
                        inserted_x=rows[n0:, 1:4].copy(), nsteps=6)
    print("b6: attempts", len(ver), "accepted", int(ver[:, 0].sum()), "max attempt no.", int(ver[:, 2].max()),
          "final atoms", len(rows))


def make_p1(wd):
    """P1: deterministic DPD forces (T = 0) with TWO atom types and per-pair a0, gamma and cutoff
    (pair_coeff i j a0 gamma rc, pair_dpd.cpp:222-251): pins the oracle's multi-type path."""
    (xlo, xhi), (ylo, yhi), (zlo, zhi) = BOX
    deck = ("units lj\nboundary f p p\natom_style atomic\natom_modify sort 0 0\ncomm_modify vel yes\n"
            "newton on\nread_data %s extra/atom/types 1\nmass 2 2.0\n"
            "set group all type/fraction 2 0.4 12345\n"
            "pair_style dpd 0.0 1.0 2616\n"
            "pair_coeff 1 1 209.6 4.5 1.0\npair_coeff 1 2 150.0 3.0 0.9\npair_coeff 2 2 100.0 6.0 0.8\n"
            "neighbor 0.4 bin\nneigh_modify delay 0 every 1\ntimestep 0.001464\n"
            "velocity all create 1.0 4928459 loop geom\n"
            "dump d all custom 1 dump.p1 id type vx vy vz fx fy fz\ndump_modify d format float %%.17g\n"
            "thermo_style custom step pe press\nthermo_modify format float %%.15g norm no\nrun 0\n" % DATA)
    lines = run_deck("p1", deck, wd)
    (_, rows), = read_dump(os.path.join(wd, "dump.p1"))
    th = None
    for k, ln in enumerate(lines):
        if ln.split()[:2] == ["Step", "PotEng"]:
            th = np.array(lines[k + 1].split(), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "p1_pair_two_types_T0.npz"), type=rows[:, 1].astype(np.int32),
                        v=rows[:, 2:5].copy(), f=rows[:, 5:8].copy(), thermo=th)
    print("p1: type-2 atoms", int((rows[:, 1] == 2).sum()), "of", len(rows), "thermo", th)


def make_b7(wd, n0):
    """B7: the B1 USHER deck with two atom types (40 % type 2, per-pair DPD coefficients) and
    `fix obmd 2 ...` (inserted atoms are type 2): pins the type-dependent branch of energy()/single()."""
    deck = (head().replace("read_data       %s" % DATA, "read_data       %s extra/atom/types 1\nmass 2 2.0\n"
                                                       "set group all type/fraction 2 0.4 12345" % DATA)
            .replace("pair_coeff      * * 209.6 4.5 1.0",
                     "pair_coeff      1 1 209.6 4.5 1.0\npair_coeff      1 2 150.0 3.0 0.9\n"
                     "pair_coeff      2 2 100.0 6.0 0.8") +
            fix(alpha=1.0, nbuf=2000).replace("fix 2 all obmd 1 1 6111", "fix 2 all obmd 2 1 6111") +
            "dump d all custom 8 dump.b7 id type x y z\ndump_modify d format float %.17g\n"
            "thermo 1\nthermo_style custom step atoms\nrun 8\n")
    assert "extra/atom/types" in deck and "pair_coeff      2 2" in deck and "obmd 2 1 6111" in deck
    lines = run_deck("b7", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b7"))
    step, rows = frames[-1]
    assert step == 8
    ver = parse_usher(lines)
    np.savez_compressed(os.path.join(HERE, "b7_usher_two_types.npz"),
                        verdicts=ver, final_tag=rows[:, 0].astype(np.int32), final_type=rows[:, 1].astype(np.int32),
                        inserted_x=rows[n0:, 2:5].copy(), nsteps=8)
    print("b7: trials", len(ver), "accepted", int(ver[:, 0].sum()), "final atoms", len(rows),
          "type-2", int((rows[:, 1] == 2).sum()))


def make_b8(wd):
    """B8: the ballistic bookkeeping deck with two atom types of mass 1 and 2 and tangential flux on
    2.5-wide shear regions: pins every mass-dependent path (FixNVE dtf/m, momentum of the deleted atoms,
    mass-weighted g_perp, fix_obmd_merged.cpp:827-837, :1344-1381, :1452-1514)."""
    deck = (head(a0=0.0, gamma=0.0, dt=0.01, shear=2.5)
            .replace("read_data       %s" % DATA, "read_data       %s extra/atom/types 1\nmass 2 2.0\n"
                                                  "set group all type/fraction 2 0.4 12345" % DATA) +
            "velocity all create 25.0 4928459 loop geom\nfix 1 all nve\n" +
            fix(alpha=1.0, tau=0.05, nbuf=2000, pxy=2.0, pxz=-1.0) +
            "dump d all custom 20 dump.b8 id type x y z vx vy vz fx fy fz\n"
            "dump_modify d format float %.17g\nrun 20\n")
    assert "extra/atom/types" in deck
    lines = run_deck("b8", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b8"))
    (s0, r0), (s1, r1) = frames
    assert (s0, s1) == (0, 20)
    ndel = sum(1 for ln in lines if ln.startswith("Deleting"))
    ver = parse_usher(lines)
    np.savez_compressed(os.path.join(HERE, "b8_ballistic_two_masses.npz"),
                        type0=r0[:, 1].astype(np.int32), v0=r0[:, 5:8].copy(),
                        tag20=r1[:, 0].astype(np.int32), type20=r1[:, 1].astype(np.int32),
                        x20=r1[:, 2:5].copy(), v20=r1[:, 5:8].copy(), f20=r1[:, 8:11].copy(),
                        ndel=ndel, nins=int(ver[:, 0].sum()))
    print("b8: deletions", ndel, "insertions", int(ver[:, 0].sum()), "atoms", len(r1))



def make_b9(wd, n0):
    """B9: the B1 USHER deck with uovlp = 150 instead of 1e4, so that the overlap branch of usher()
    (fix_obmd_merged.cpp:1566-1580: ds = dsovlp - pow(4 eps / E, 1/12)) is taken by the many trial
    points whose energy exceeds 150 (the mean energy of a random point in this fluid is ~132)."""
    deck = (head() + fix(alpha=1.0, nbuf=2000).replace("usher 1 31.03 1.0 0.02 10000 1.5 1.0 40",
                                                       "usher 1 31.03 1.0 0.02 150 1.5 1.0 40") +
            "dump d all custom 10 dump.b9 id x y z\ndump_modify d format float %.17g\n"
            "thermo 1\nthermo_style custom step atoms\nrun 10\n")
    assert "0.02 150 1.5" in deck
    lines = run_deck("b9", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b9"))
    step, rows = frames[-1]
    assert step == 10
    ver = parse_usher(lines)
    np.savez_compressed(os.path.join(HERE, "b9_usher_overlap.npz"),
                        verdicts=ver, final_tag=rows[:, 0].astype(np.int32),
                        inserted_x=rows[n0:, 1:4].copy(), nsteps=10)
    print("b9: trials", len(ver), "accepted", int(ver[:, 0].sum()), "final atoms", len(rows))


def make_m0(wd):
    """M0: the bonded path of BASELINE config 4 in miniature.  The shipped atoms, rewritten
    in atom_style bond: nearest neighbours closer than 0.9 are paired into dimers (bond type 1), and every
    third dimer is linked to the next free atom within 1.1 (bond type 2) -- ~4000 bonds, some of them across
    the periodic y / z faces.  `bond_style harmonic` (MOLECULE/bond_harmonic.cpp:48-103), `special_bonds lj
    1 1 1` (no pair exclusions), T = 0.  Two runs of the reference binary: bonds only (`pair_style zero`)
    and bonds + DPD; forces, ebond."""
    from scipy.spatial import cKDTree
    rows = [ln.split() for ln in open(DATA).read().split("Atoms")[1].strip().split("\n") if len(ln.split()) == 5]
    tag = np.array([int(r[0]) for r in rows]); x = np.array([[float(v) for v in r[2:5]] for r in rows])
    (xlo, xhi), (ylo, yhi), (zlo, zhi) = BOX
    L = np.array([1.0e6, yhi - ylo, zhi - zlo])          # x is not periodic
    xw = x.copy(); xw[:, 1:] = np.mod(xw[:, 1:], L[1:])
    tree = cKDTree(xw, boxsize=L)
    free = np.ones(len(x), bool)
    bonds = []
    pairs = sorted(tree.query_pairs(0.9))
    ndim = 0
    for i, j in pairs:
        if free[i] and free[j]:
            free[i] = free[j] = False
            bonds.append((1, tag[i], tag[j]))
            ndim += 1
            if ndim % 3 == 0:
                for k in tree.query_ball_point(xw[j], 1.1):
                    if free[k]:
                        free[k] = False
                        bonds.append((2, tag[j], tag[k]))
                        break
    bonds = np.array(bonds, dtype=np.int32)
    mol = np.zeros(len(x), np.int32)
    data = os.path.join(wd, "m0.data")
    with open(data, "w") as fh:
        fh.write("bonded DPD beads\n\n%d atoms\n1 atom types\n%d bonds\n2 bond types\n\n" % (len(x), len(bonds)))
        for (lo, hi), n in zip(BOX, "xyz"):
            fh.write("%.17g %.17g %slo %shi\n" % (lo, hi, n, n))
        fh.write("\nMasses\n\n1 1\n\nAtoms # bond\n\n")
        for t, m, r in zip(tag, mol, x):
            fh.write("%d %d 1 %.17g %.17g %.17g\n" % (t, m, r[0], r[1], r[2]))
        fh.write("\nBonds\n\n")
        for k, (bt, a, b) in enumerate(bonds):
            fh.write("%d %d %d %d\n" % (k + 1, bt, a, b))
    styles = {"m0": ("bond_style harmonic\nbond_coeff 1 100.0 0.5\nbond_coeff 2 50.0 0.8\n", "m0_bond_harmonic.npz",
                     dict(k=np.array([0.0, 100.0, 50.0]), r0=np.array([0.0, 0.5, 0.8]))),
              # bond_style fene (MOLECULE/bond_fene.cpp:53-135): K R0 epsilon sigma; the WCA part acts below 2^(1/6) sigma
              "m1": ("bond_style fene\nbond_coeff 1 30.0 1.5 1.0 0.6\nbond_coeff 2 20.0 1.8 0.5 0.7\n", "m1_bond_fene.npz",
                     dict(k=np.array([0.0, 30.0, 20.0]), r0=np.array([0.0, 1.5, 1.8]), epsilon=np.array([0.0, 1.0, 0.5]),
                          sigma=np.array([0.0, 0.6, 0.7])))}
    for key, (bond_cmds, fname, coeffs) in styles.items():
        out = {}
        for name, pair in (("bond", "pair_style zero 1.0\npair_coeff * *\n"),
                           ("both", "pair_style dpd 0.0 1.0 2616\npair_coeff * * 209.6 4.5 1.0\n")):
            deck = ("units lj\nboundary f p p\natom_style bond\natom_modify sort 0 0\ncomm_modify vel yes\n"
                    "newton on\nread_data %s\n%s%s"
                    "special_bonds lj 1.0 1.0 1.0\nneighbor 0.4 bin\nneigh_modify delay 0 every 1\ntimestep 0.001464\n"
                    "velocity all create 1.0 4928459 loop geom\n"
                    "dump d all custom 1 dump.m0 id vx vy vz fx fy fz\ndump_modify d format float %%.17g\n"
                    "compute pb all pressure NULL bond\n"
                    "thermo_style custom step pe ebond c_pb[1] c_pb[2] c_pb[3] c_pb[4] c_pb[5] c_pb[6] vol\n"
                    "thermo_modify format float %%.15g norm no\nrun 0\n" % (data, pair, bond_cmds))
            lines = run_deck(key + name, deck, wd)
            (_, r), = read_dump(os.path.join(wd, "dump.m0"))
            th = None
            for k, ln in enumerate(lines):
                if ln.split()[:2] == ["Step", "PotEng"]:
                    th = np.array(lines[k + 1].split(), dtype=np.float64)
            assert np.array_equal(r[:, 0].astype(np.int64), tag)
            out[name] = (r, th)
        np.savez_compressed(os.path.join(HERE, fname), bonds=bonds, v=out["both"][0][:, 1:4].copy(),
                            f_bond=out["bond"][0][:, 4:7].copy(), f_both=out["both"][0][:, 4:7].copy(),
                            ebond=out["bond"][1][2], thermo_both=out["both"][1],
                            # virial of the bond forces: c_pb = W_ab / V (compute pressure NULL bond; xx yy zz xy xz yz)
                            virial_bond=out["bond"][1][3:9] * out["bond"][1][9], **coeffs)
        print(key, ": bonds", len(bonds), "type-2", int((bonds[:, 0] == 2).sum()), "ebond", out["bond"][1][2],
              "|f_bond|max", np.abs(out["bond"][0][:, 4:7]).max())


def read_chunk_file(path, ncol):
    """fix ave/chunk text output -> list of (step, array[nchunk, ncol])."""
    out = []
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("#")]
    i = 0
    while i < len(lines):
        w = lines[i].split()
        step, nch = int(w[0]), int(w[1])
        rows = np.array([ln.split() for ln in lines[i + 1:i + 1 + nch]], dtype=np.float64)
        out.append((step, rows[:, :ncol].copy()))
        i += 1 + nch
    return out


def make_r1_long(wd, equil=2000, sample=2000, block=200, nbins=48):
    """Steady-exchange regime R1 of the shipped deck: `equil` steps from v = 0, a snapshot
    of x and v (the start state bench.py tiles for BOTH arms), then `sample` more steps with the 1-D
    profiles along x that the statistical parity test compares (compute chunk/atom + fix ave/chunk:
    number density, <vx>, kinetic temperature, per-atom stress xx), one row per block of `block`
    steps so the test can form block-averaged error bars."""
    deck = (head().replace("atom_modify     sort 0 0\n", "atom_modify     sort 0 0\n") + "fix 1 all nve\n" +
            fix(alpha=1.0, nbuf=1895) +
            "thermo 100\nthermo_style custom step atoms temp press\n"
            "thermo_modify format float %%.15g lost ignore\n"
            "run %d\n"
            "write_dump all custom dump.snap id x y z vx vy vz modify format float %%.17g\n"
            "compute cc all chunk/atom bin/1d x lower %.17g units box\n"
            "compute st all stress/atom NULL\n"
            "fix av all ave/chunk 1 %d %d cc density/number vx temp c_st[1] c_st[2] c_st[3] norm all file prof.txt\n"
            "run %d\n" % (equil, BOX[0][1] / nbins, block, block, sample))
    lines = run_deck("r1long", deck, wd)
    (_, rows), = read_dump(os.path.join(wd, "dump.snap"))
    order = np.argsort(rows[:, 0], kind="stable")
    rows = rows[order]
    for dst in (HERE, os.path.join(ROOT, "obmd-lammps-extension_b200", "data")):   # the second copy feeds bench.py
        os.makedirs(dst, exist_ok=True)
        np.savez_compressed(os.path.join(dst, "equil_r1.npz"), tag=rows[:, 0].astype(np.int32),
                            x=rows[:, 1:4].copy(), v=rows[:, 4:7].copy(), box=np.array(BOX), steps=equil)
    # columns of fix ave/chunk: chunk coord ncount density/number vx temp c_st[1..3]
    prof = read_chunk_file(os.path.join(wd, "prof.txt"), 9)
    th = []
    on = False
    for ln in lines:
        w = ln.split()
        if w[:2] == ["Step", "Atoms"]:
            on = True; continue
        if on:
            if len(w) == 4 and w[0].isdigit():
                th.append([float(t) for t in w])
            elif w and w[0] == "Loop":
                on = False
    ver = parse_usher(lines)
    ndel = sum(1 for ln in lines if ln.startswith("Deleting"))
    np.savez_compressed(os.path.join(HERE, "r1_profiles.npz"), blocks=np.array([r for _, r in prof]),
                        block_steps=np.array([s for s, _ in prof]), thermo=np.array(th), nbins=nbins,
                        equil=equil, sample=sample, block=block, ndel=ndel, ntrials=len(ver),
                        nacc=int(ver[:, 0].sum()))
    print("r1_long: snapshot atoms", len(rows), "blocks", len(prof), "trials", len(ver),
          "accepted", int(ver[:, 0].sum()), "deleted", ndel)


def main():
    if not os.path.exists(LMP):
        sys.exit("oracle/_ref/lmp_serial missing: run `bash oracle/build_ref.sh` first")
    wd = tempfile.mkdtemp(prefix="obmd_golden_")
    print("working in", wd)
    if len(sys.argv) > 1 and sys.argv[1] == "m0":   # only the bonded-path fixture
        make_m0(wd)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "b6":   # only the maxattempt fixture (the others are unchanged)
        make_b6(wd, 12637)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "p1":   # only the two-type pair fixture
        make_p1(wd)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "b8":   # only the two-mass ballistic fixture
        make_b8(wd)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "b9":   # only the overlap-branch USHER fixture
        make_b9(wd, 12637)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "r1long":   # only the long R1 run (snapshot + profiles), ~20 min
        make_r1_long(wd)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "b7":   # only the two-type USHER fixture
        make_b7(wd, 12637)
        return

    # ---- shipped configuration (input fixture: ids + coordinates, read_data order) ----
    deck = head() + "dump d all custom 1 dump.in id type x y z\ndump_modify d format float %.17g\nrun 0\n"
    run_deck("input", deck, wd)
    (_, rows), = read_dump(os.path.join(wd, "dump.in"))
    tag0 = rows[:, 0].astype(np.int32); x0 = rows[:, 2:5].copy()
    box = np.array(BOX)
    np.savez_compressed(os.path.join(HERE, "dpd_8map_obmd.npz"), tag=tag0, x=x0, box=box)
    print("input:", len(tag0), "atoms")

    # ---- B1: USHER / RanPark / tag stream on static atoms ----
    deck = (head() + fix(alpha=1.0, nbuf=2000) +
            "dump d all custom 12 dump.b1 id x y z\ndump_modify d format float %.17g\n"
            "thermo 1\nthermo_style custom step atoms\nrun 12\n")
    lines = run_deck("b1", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b1"))
    step, rows = frames[-1]
    assert step == 12
    ver = parse_usher(lines)
    n0 = len(tag0)
    np.savez_compressed(os.path.join(HERE, "b1_usher_static.npz"),
                        verdicts=ver, final_tag=rows[:, 0].astype(np.int32),
                        inserted_x=rows[n0:, 1:4].copy(), nsteps=12)
    print("b1: trials", len(ver), "accepted", int(ver[:, 0].sum()), "final atoms", len(rows))
    make_b6(wd, n0)
    make_b7(wd, n0)
    make_b8(wd)
    make_b9(wd, n0)

    # ---- B2/B5: normal force distribution, static atoms, pair force zeroed ----
    deck = (head(a0=0.0, gamma=0.0) + fix(dpxx=10.0, freq=5.0) +
            "dump d all custom 1 dump.b2 id fx fy fz\ndump_modify d format float %.17g\nrun 3\n")
    run_deck("b2", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b2"))
    np.savez_compressed(os.path.join(HERE, "b2_normal_force.npz"),
                        **{"f%d" % s: r[:, 1:4].copy() for s, r in frames})
    print("b2: frames", [s for s, _ in frames])

    # ---- B4: tangential force, shear regions 2.5 wide ----
    deck = (head(a0=0.0, gamma=0.0, shear=2.5) + fix(pxy=2.0, pxz=-1.0) +
            "dump d all custom 1 dump.b4 id fx fy fz\ndump_modify d format float %.17g\nrun 1\n")
    run_deck("b4", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b4"))
    np.savez_compressed(os.path.join(HERE, "b4_shear_force.npz"), f1=frames[-1][1][:, 1:4].copy())
    print("b4: ok")

    # ---- B3: ballistic deletion / insertion bookkeeping, 20 steps ----
    deck = (head(a0=0.0, gamma=0.0, dt=0.01) +
            "velocity all create 25.0 4928459 loop geom\nfix 1 all nve\n" +
            fix(alpha=1.0, tau=0.05, nbuf=2000) +
            "dump d all custom 20 dump.b3 id x y z vx vy vz fx fy fz\n"
            "dump_modify d format float %.17g\nrun 20\n")
    lines = run_deck("b3", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.b3"))
    (s0, r0), (s1, r1) = frames
    assert (s0, s1) == (0, 20)
    ndel = sum(1 for ln in lines if ln.startswith("Deleting"))
    ver = parse_usher(lines)
    np.savez_compressed(os.path.join(HERE, "b3_ballistic.npz"),
                        v0=r0[:, 4:7].copy(), tag20=r1[:, 0].astype(np.int32),
                        x20=r1[:, 1:4].copy(), v20=r1[:, 4:7].copy(), f20=r1[:, 7:10].copy(),
                        ndel=ndel, nins=int(ver[:, 0].sum()))
    print("b3: deletions", ndel, "insertions", int(ver[:, 0].sum()), "atoms", len(r1))

    # ---- deterministic DPD forces: T = 0 (sigma = 0), thermal velocities ----
    deck = (head(T=0.0) + "velocity all create 1.0 4928459 loop geom\n"
            "dump d all custom 1 dump.p0 id vx vy vz fx fy fz\ndump_modify d format float %.17g\n"
            "thermo_style custom step pe press pxx pyy pzz pxy pxz pyz\n"
            "thermo_modify format float %.15g norm no\nrun 0\n")
    lines = run_deck("p0", deck, wd)
    (_, rows), = read_dump(os.path.join(wd, "dump.p0"))
    th = None
    for k, ln in enumerate(lines):
        if ln.split()[:2] == ["Step", "PotEng"]:
            th = np.array(lines[k + 1].split(), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "p0_pair_T0.npz"), v=rows[:, 1:4].copy(),
                        f=rows[:, 4:7].copy(), thermo=th)
    print("p0: thermo", th)
    make_p1(wd)

    # ---- RanMars: two static atoms, one Gaussian per step ----
    two = os.path.join(wd, "two.data")
    with open(two, "w") as fh:
        fh.write("two atoms\n\n2 atoms\n1 atom types\n\n0.0 10.0 xlo xhi\n0.0 10.0 ylo yhi\n"
                 "0.0 10.0 zlo zhi\n\nMasses\n\n1 1.0\n\nAtoms\n\n1 1 5.0 5.0 5.0\n2 1 5.5 5.0 5.0\n")
    deck = ("units lj\nboundary f p p\natom_style atomic\natom_modify sort 0 0\ncomm_modify vel yes\n"
            "read_data two.data\npair_style dpd 1.0 1.0 2616\npair_coeff * * 209.6 4.5 1.0\n"
            "timestep 0.001464\ndump d all custom 1 dump.rm id fx\n"
            "dump_modify d format float %.17g\nrun 16\n")
    run_deck("rm", deck, wd)
    frames = read_dump(os.path.join(wd, "dump.rm"))
    fx = np.array([r[0, 1] for _, r in frames])
    np.savez_compressed(os.path.join(HERE, "ranmars_two_atoms.npz"), fx=fx)
    print("ranmars: fx[0..2]", fx[:3])

    # ---- R1 steady-exchange thermo trace (statistical parity reference) ----
    deck = (head() + "fix 1 all nve\n" + fix(alpha=1.0, nbuf=1895) +
            "thermo 10\nthermo_style custom step atoms temp press\n"
            "thermo_modify format float %.15g lost ignore\nrun 400\n")
    lines = run_deck("r1", deck, wd)
    th = []
    on = False
    for ln in lines:
        w = ln.split()
        if w[:2] == ["Step", "Atoms"]:
            on = True; continue
        if on:
            if len(w) == 4 and w[0].isdigit():
                th.append([float(t) for t in w])
            elif w and w[0] == "Loop":
                on = False
    ver = parse_usher(lines)
    ndel = sum(1 for ln in lines if ln.startswith("Deleting"))
    np.savez_compressed(os.path.join(HERE, "r1_thermo.npz"), thermo=np.array(th), ndel=ndel,
                        ntrials=len(ver), nacc=int(ver[:, 0].sum()))
    print("r1: steps", len(th), "ndel", ndel, "trials", len(ver), "accepted", int(ver[:, 0].sum()))


if __name__ == "__main__":
    main()
