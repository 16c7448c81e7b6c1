"""Drop-in test: the UNMODIFIED reference LAMMPS binary (oracle/_ref/lmp_serial) loads
obmd-lammps-extension_b200/libobmd_b200_lammps.so with `plugin load`, which overrides the built-in
`fix obmd` and `pair_style dpd`; the same decks that produced the golden fixtures are then run
through the CUDA path and compared with those fixtures.  GPU box only."""
import os
import subprocess

import numpy as np
import pytest

import common_cases as cc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LMP = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
PLUGIN = os.path.join(ROOT, "obmd-lammps-extension_b200", "libobmd_b200_lammps.so")

HEAD = """plugin load {plugin}
units           lj
boundary        f p p
atom_style      atomic
atom_modify     sort 0 0
comm_modify     vel yes
newton          on
read_data       {data}
region          leftB block {xlo:.17g} {xlB:.17g} {ylo:.17g} {yhi:.17g} {zlo:.17g} {zhi:.17g}
region          rightB block {xrB:.17g} {xhi:.17g} {ylo:.17g} {yhi:.17g} {zlo:.17g} {zhi:.17g}
region          leftshear block 0.0 0.0 0.0 0.0 0.0 0.0
region          rightshear block 0.0 0.0 0.0 0.0 0.0 0.0
pair_style      dpd {T} 1.0 2616
pair_coeff      * * {a0} {gamma} 1.0
neighbor        0.4 bin
neigh_modify    delay 0 every 1
timestep        {dt}
"""
FIX = ("fix 2 all obmd 1 1 6111 188.0 0 0 0 0 {alpha} {tau} {nbuf} region1 leftB region2 rightB "
       "region3 leftshear region4 rightshear region5 leftB region6 rightB buffersize {B:.17g} gfac 0.25 "
       "stepparallel 0 stepperp 1 maxattempt 1 usher 1 31.03 1.0 0.02 10000 1.5 1.0 40 charged 0\n")


def need():
    if not (os.path.exists(LMP) and os.path.exists(PLUGIN)):
        pytest.skip("reference binary or plugin not built (oracle/build_ref.sh, host/Makefile)")


def write_data(path, tag, x, box):
    with open(path, "w") as fh:
        fh.write("shipped OBMD_DPD configuration\n\n%d atoms\n1 atom types\n\n" % len(x))
        for d, n in enumerate("xyz"):
            fh.write("%.17g %.17g %slo %shi\n" % (box[d, 0], box[d, 1], n, n))
        fh.write("\nMasses\n\n1 1\n\nAtoms\n\n")
        for t, r in zip(tag, x):
            fh.write("%d 1 %.17g %.17g %.17g\n" % (t, r[0], r[1], r[2]))


def read_dump(path):
    frames = []
    lines = open(path).read().split("\n")
    i = 0
    while i < len(lines):
        if lines[i].startswith("ITEM: TIMESTEP"):
            step = int(lines[i + 1]); n = int(lines[i + 3])
            rows = np.array([ln.split() for ln in lines[i + 9:i + 9 + n]], dtype=np.float64)
            frames.append((step, rows)); i += 9 + n
        else:
            i += 1
    return frames


def run(tmp_path, body, **kw):
    tag, x, box = cc.shipped()
    data = str(tmp_path / "shipped.data")
    write_data(data, tag, x, box)
    B = 0.15 * box[0, 1]
    d = dict(plugin=PLUGIN, data=data, T=1.0, a0=209.6, gamma=4.5, dt=0.001464, xlo=box[0, 0], xhi=box[0, 1],
             ylo=box[1, 0], yhi=box[1, 1], zlo=box[2, 0], zhi=box[2, 1], xlB=box[0, 0] + B, xrB=box[0, 1] - B)
    d.update(kw)
    deck = HEAD.format(**d) + body
    inp = str(tmp_path / "in.deck")
    open(inp, "w").write(deck)
    out = subprocess.run([LMP, "-in", inp, "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "Overriding built-in fix style obmd from plugin" in out.stdout + out.stderr or \
        "Overriding built-in pair style dpd from plugin" in out.stdout + out.stderr
    return out.stdout, tag, x, box, B


def test_plugin_pair_forces_T0(tmp_path):
    need()
    body = ("velocity all create 1.0 4928459 loop geom\n"
            "dump d all custom 1 dump.p0 id vx vy vz fx fy fz\ndump_modify d format float %.17g\n"
            "thermo_style custom step pe press pxx pyy pzz pxy pxz pyz\n"
            "thermo_modify format float %.15g norm no\nrun 0\n")
    out, tag, x, box, B = run(tmp_path, body, T=0.0)
    g = cc.load("p0_pair_T0.npz")
    (_, rows), = read_dump(str(tmp_path / "dump.p0"))
    assert np.array_equal(rows[:, 1:4], g["v"])            # same LAMMPS, same velocity command
    assert cc.mixed_rel_err(rows[:, 4:7], g["f"]) < 1e-9
    th = None
    lines = out.split("\n")
    for k, ln in enumerate(lines):
        if ln.split()[:2] == ["Step", "PotEng"]:
            th = np.array(lines[k + 1].split(), dtype=np.float64)
    assert th is not None and np.allclose(th[1:], g["thermo"][1:], rtol=1e-9, atol=1e-9)


def test_plugin_usher_static_bit_exact(tmp_path):
    need()
    tag, x, box = cc.shipped()
    B = 0.15 * box[0, 1]
    body = (FIX.format(alpha=1.0, tau=0.005, nbuf=2000, B=B) +
            "dump d all custom 12 dump.b1 id x y z\ndump_modify d format float %.17g\nrun 12\n")
    run(tmp_path, body)
    g = cc.load("b1_usher_static.npz")
    step, rows = read_dump(str(tmp_path / "dump.b1"))[-1]
    assert step == 12
    assert np.array_equal(rows[:, 0].astype(np.int32), g["final_tag"])
    assert np.array_equal(rows[len(tag):, 1:4], g["inserted_x"])


def test_plugin_ballistic_bookkeeping(tmp_path):
    need()
    tag, x, box = cc.shipped()
    B = 0.15 * box[0, 1]
    body = ("velocity all create 25.0 4928459 loop geom\nfix 1 all nve\n" +
            FIX.format(alpha=1.0, tau=0.05, nbuf=2000, B=B) +
            "dump d all custom 20 dump.b3 id x y z vx vy vz fx fy fz\n"
            "dump_modify d format float %.17g\nrun 20\n")
    run(tmp_path, body, a0=0.0, gamma=0.0, dt=0.01)
    g = cc.load("b3_ballistic.npz")
    (s0, r0), (s1, r1) = read_dump(str(tmp_path / "dump.b3"))
    assert (s0, s1) == (0, 20)
    assert np.array_equal(r1[:, 0].astype(np.int32), g["tag20"])
    assert np.allclose(r1[:, 1:4], g["x20"], rtol=0, atol=1e-11)
    assert np.allclose(r1[:, 4:7], g["v20"], rtol=0, atol=1e-10)
    assert cc.mixed_rel_err(r1[:, 7:10], g["f20"]) < 1e-9


def test_plugin_run_obmd_device_resident_ballistic(tmp_path):
    """Stage B: `run_obmd 20` (command style of the same plugin) runs the b3 deck device-resident
    (obmd_run) and brings the atoms back only for the dump at step 20: identical tags, x, v, f vs the
    fixture the unmodified reference produced with its own `run 20`."""
    need()
    tag, x, box = cc.shipped()
    B = 0.15 * box[0, 1]
    body = ("velocity all create 25.0 4928459 loop geom\nfix 1 all nve\n" +
            FIX.format(alpha=1.0, tau=0.05, nbuf=2000, B=B) +
            "dump d all custom 20 dump.b3 id x y z vx vy vz fx fy fz\n"
            "dump_modify d format float %.17g\nthermo 10\nthermo_style custom step atoms temp pe press\n"
            "run_obmd 20\n")
    out, *_ = run(tmp_path, body, a0=0.0, gamma=0.0, dt=0.01)
    g = cc.load("b3_ballistic.npz")
    (s0, r0), (s1, r1) = read_dump(str(tmp_path / "dump.b3"))
    assert (s0, s1) == (0, 20)
    assert np.array_equal(r1[:, 0].astype(np.int32), g["tag20"])
    assert np.allclose(r1[:, 1:4], g["x20"], rtol=0, atol=1e-11)
    assert np.allclose(r1[:, 4:7], g["v20"], rtol=0, atol=1e-10)
    assert cc.mixed_rel_err(r1[:, 7:10], g["f20"]) < 1e-9
    assert "Loop time of" in out
    # thermo rows at steps 0, 10, 20 with the atom count the fix keeps (natoms)
    rows = [ln.split() for ln in out.split("\n") if len(ln.split()) == 5 and ln.split()[0] in ("0", "10", "20")]
    assert [r[0] for r in rows] == ["0", "10", "20"] and int(rows[-1][1]) == len(r1)


def test_plugin_run_obmd_steady_exchange_and_restart(tmp_path):
    """Stage B on the real deck (T = 1, USHER active), then write_restart / read_restart: the fix state
    (RanPark seed, simulation time, counters: obmd_fix_get/set_state) survives, so a run continued from
    the restart file inserts with the same RanPark stream as the uninterrupted run."""
    need()
    tag, x, box = cc.shipped()
    B = 0.15 * box[0, 1]
    fix = FIX.format(alpha=1.0, tau=0.005, nbuf=1895, B=B)
    body = ("fix 1 all nve\n" + fix + "thermo 50\nthermo_style custom step atoms temp\n"
            "run_obmd 100\nwrite_restart obmd.restart\n"
            "dump d all custom 10 dump.a id x y z\ndump_modify d format float %.17g\nrun_obmd 10\n")
    out, *_ = run(tmp_path, body)
    (_, ra), = [f for f in read_dump(str(tmp_path / "dump.a")) if f[0] == 110]
    # continue from the restart file in a fresh LAMMPS
    deck = ("plugin load %s\nread_restart obmd.restart\ncomm_modify vel yes\natom_modify sort 0 0\n"
            "region leftB block %.17g %.17g %.17g %.17g %.17g %.17g\n"
            "region rightB block %.17g %.17g %.17g %.17g %.17g %.17g\n"
            "region leftshear block 0.0 0.0 0.0 0.0 0.0 0.0\nregion rightshear block 0.0 0.0 0.0 0.0 0.0 0.0\n"
            "pair_style dpd 1.0 1.0 2616\npair_coeff * * 209.6 4.5 1.0\nneighbor 0.4 bin\ntimestep 0.001464\n"
            % (PLUGIN, box[0, 0], box[0, 0] + B, box[1, 0], box[1, 1], box[2, 0], box[2, 1],
               box[0, 1] - B, box[0, 1], box[1, 0], box[1, 1], box[2, 0], box[2, 1]) +
            "fix 1 all nve\n" + fix +
            "dump d all custom 10 dump.b id x y z\ndump_modify d format float %.17g\nrun_obmd 10\n")
    inp = str(tmp_path / "in.restart")
    open(inp, "w").write(deck)
    r = subprocess.run([LMP, "-in", inp, "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    (_, rb), = [f for f in read_dump(str(tmp_path / "dump.b")) if f[0] == 110]
    # same atoms with the same tags (~70 insertions drew the same RanPark stream); positions agree to the
    # accuracy the restart allows: forces are recomputed at setup in another atom order, and the dense
    # fluid amplifies that rounding difference by ~10 orders of magnitude in 50 steps (which is why the
    # comparison is made 10 steps after the restart, not 50: by then a USHER verdict may flip)
    assert np.array_equal(np.sort(ra[:, 0]), np.sort(rb[:, 0]))
    oa, ob_ = np.argsort(ra[:, 0]), np.argsort(rb[:, 0])
    assert np.abs(ra[oa, 1:4] - rb[ob_, 1:4]).max() < 1e-9
    assert len(ra) != len(tag)      # the run did insert / delete


def run_expecting_error(tmp_path, body, head=HEAD, **kw):
    tag, x, box = cc.shipped()
    data = str(tmp_path / "shipped.data")
    write_data(data, tag, x, box)
    B = 0.15 * box[0, 1]
    d = dict(plugin=PLUGIN, data=data, T=1.0, a0=209.6, gamma=4.5, dt=0.001464, xlo=box[0, 0], xhi=box[0, 1],
             ylo=box[1, 0], yhi=box[1, 1], zlo=box[2, 0], zhi=box[2, 1], xlB=box[0, 0] + B, xrB=box[0, 1] - B)
    d.update(kw)
    inp = str(tmp_path / "in.deck")
    open(inp, "w").write(head.format(**d) + body)
    out = subprocess.run([LMP, "-in", inp, "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True)
    assert out.returncode != 0, "LAMMPS accepted the deck:\n" + out.stdout[-2000:]
    return out.stdout + out.stderr, B


def test_plugin_argument_errors(tmp_path):
    """The C++ argument parser that ships (host/lammps_plugin.cpp: FixObmdB200 constructor / options,
    PairDPDB200::settings / coeff), driven through the reference binary: malformed commands stop LAMMPS
    with the reference's own messages (fix_obmd_merged.cpp:104-135, :1990-2130; pair_dpd.cpp:195-251)."""
    need()
    B = 0.15 * cc.shipped()[2][0, 1]
    good = FIX.format(alpha=0.7, tau=0.005, nbuf=1327, B=B)
    cases = [
        ("fix 2 all obmd 1 1 6111 188.0 0 0 0 0 0.7 0.005\n", "Illegal fix obmd command (check # of arguments)"),
        (good.replace(" 6111 ", " 0 "), "Illegal fix obmd command"),
        (good.replace("obmd 1 1", "obmd 7 1"), "Invalid atom type in fix obmd command"),
        (good.replace("region1 leftB", "region1 nosuch"), "for fix obmd does not exist"),
        (good.replace("gfac 0.25", "gfac 1.5"), "Parameter gfac should be between 0.0 and 1.0"),
        (good.replace("buffersize %.17g" % B, "buffersize -1.0"), "Parameter buffersize should be larger than 0.0"),
        (good.replace("usher 1 31.03", "usher 0 31.03"), "You must pick either usher or near"),
        (good.replace("stepparallel 0", "stepparallel 1"), "only the smooth transition distribution function"),
        (good.replace("charged 0", "bogus 0"), "Illegal fix obmd command"),
        (good.replace("region1 leftB ", ""), "Must specify a region in fix obmd"),
        (good.replace("region5 leftB ", ""), "Region ID for fix obmd does not exist"),
    ]
    for k, (fixline, msg) in enumerate(cases):
        sub = tmp_path / ("case%d" % k)
        sub.mkdir()
        out, _ = run_expecting_error(sub, "fix 1 all nve\n" + fixline + "run 0\n")
        assert msg in out, (fixline, out[-1500:])
    # pair_style / pair_coeff (pair_dpd.cpp:195-251)
    bad_style = HEAD.replace("pair_style      dpd {T} 1.0 2616", "pair_style      dpd {T} 1.0")
    out, _ = run_expecting_error(tmp_path, "run 0\n", head=bad_style)
    assert "Illegal pair_style command" in out
    bad_seed = HEAD.replace("pair_style      dpd {T} 1.0 2616", "pair_style      dpd {T} 1.0 0")
    out, _ = run_expecting_error(tmp_path, "run 0\n", head=bad_seed)
    assert "Illegal pair_style command" in out
    bad_coeff = HEAD.replace("pair_coeff      * * {a0} {gamma} 1.0", "pair_coeff      * * {a0}")
    out, _ = run_expecting_error(tmp_path, "run 0\n", head=bad_coeff)
    assert "Incorrect args for pair coefficients" in out
    # the working deck still needs `atom_modify sort 0 0` (the device keeps the LAMMPS local order)
    no_sort = HEAD.replace("atom_modify     sort 0 0\n", "")
    out, _ = run_expecting_error(tmp_path, "fix 1 all nve\n" + good + "run 0\n", head=no_sort)
    assert "requires 'atom_modify sort 0 0'" in out


def test_plugin_bond_harmonic(tmp_path):
    """`bond_style harmonic` overridden by the plugin (BondHarmonicB200 -> obmd_bond_compute) next to its
    `pair_style dpd`, in atom_style bond inside the reference binary: forces, bond energy and bond virial
    of the fixture deck (tests/golden/make_golden.py m0: 6803 bonds of two types) against what the
    built-in styles produced."""
    need()
    tag, x, box = cc.shipped()
    g = cc.load("m0_bond_harmonic.npz")
    b = g["bonds"]
    data = str(tmp_path / "m0.data")
    with open(data, "w") as fh:
        fh.write("bonded DPD beads\n\n%d atoms\n1 atom types\n%d bonds\n2 bond types\n\n" % (len(x), len(b)))
        for d, n in enumerate("xyz"):
            fh.write("%.17g %.17g %slo %shi\n" % (box[d, 0], box[d, 1], n, n))
        fh.write("\nMasses\n\n1 1\n\nAtoms # bond\n\n")
        for t, r in zip(tag, x):
            fh.write("%d 0 1 %.17g %.17g %.17g\n" % (t, r[0], r[1], r[2]))
        fh.write("\nBonds\n\n")
        for k, (bt, a1, a2) in enumerate(b):
            fh.write("%d %d %d %d\n" % (k + 1, bt, a1, a2))
    deck = ("plugin load %s\nunits lj\nboundary f p p\natom_style bond\natom_modify sort 0 0\ncomm_modify vel yes\n"
            "newton on\nread_data %s\npair_style dpd 0.0 1.0 2616\npair_coeff * * 209.6 4.5 1.0\n"
            "bond_style harmonic\nbond_coeff 1 100.0 0.5\nbond_coeff 2 50.0 0.8\nspecial_bonds lj 1.0 1.0 1.0\n"
            "neighbor 0.4 bin\nneigh_modify delay 0 every 1\ntimestep 0.001464\n"
            "velocity all create 1.0 4928459 loop geom\n"
            "dump d all custom 1 dump.m0 id vx vy vz fx fy fz\ndump_modify d format float %%.17g\n"
            "compute pb all pressure NULL bond\n"
            "thermo_style custom step pe ebond c_pb[1] c_pb[2] c_pb[3] c_pb[4] c_pb[5] c_pb[6] vol\n"
            "thermo_modify format float %%.15g norm no\nrun 0\n" % (PLUGIN, data))
    inp = str(tmp_path / "in.m0")
    open(inp, "w").write(deck)
    r = subprocess.run([LMP, "-in", inp, "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    out = r.stdout + r.stderr
    assert "Overriding built-in bond style harmonic from plugin" in out
    (_, rows), = read_dump(str(tmp_path / "dump.m0"))
    assert np.array_equal(rows[:, 1:4], g["v"])
    assert np.abs(rows[:, 4:7] - g["f_both"]).max() <= 1e-9 * np.abs(g["f_both"]).max()
    th = None
    lines = r.stdout.split("\n")
    for k, ln in enumerate(lines):
        if ln.split()[:2] == ["Step", "PotEng"]:
            th = np.array(lines[k + 1].split(), dtype=np.float64)
    assert th is not None
    assert abs(th[2] - float(g["ebond"])) <= 1e-10 * abs(float(g["ebond"]))
    w = th[3:9] * th[9]
    assert np.abs(w - g["virial_bond"]).max() <= 1e-9 * np.abs(g["virial_bond"]).max()
    assert abs(th[1] - g["thermo_both"][1]) <= 1e-9 * abs(g["thermo_both"][1])      # pe = evdwl + ebond
