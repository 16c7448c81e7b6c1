"""Host-side logic of the multi-GPU path on CPU: slab partition helpers and the rendezvous
plumbing (unique-id broadcast, packed count reduction) under world_size-2 gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_is_complete_and_disjoint():
    import obmd_b200 as ob
    for ncols in (7, 67, 134, 671):
        for P in (1, 2, 3, 4, 8):
            if ncols < 2 * P:
                continue
            cover = []
            for r in range(P):
                c0, c1 = ob.slab_columns(ncols, P, r)
                assert c1 - c0 >= ncols // P
                cover += list(range(c0, c1))
            assert cover == list(range(ncols))
    # non-uniform slabs (BASELINE config 5: 16 M atoms, 672 columns, buffer 0.15 Lx = 101 columns + margin > 672 / 8)
    for ncols, P, edge in ((672, 8, 103), (672, 4, 103), (167, 4, 60), (40, 8, 9)):
        cover, widths = [], []
        for r in range(P):
            c0, c1 = ob.slab_columns(ncols, P, r, edge)
            cover += list(range(c0, c1)); widths.append(c1 - c0)
        assert cover == list(range(ncols))
        if edge > ncols // P:
            assert widths[0] == widths[-1] == edge
            assert max(widths[1:-1]) - min(widths[1:-1]) <= 1
    with pytest.raises(ob.ObmdError):
        ob.slab_columns(40, 8, 3, 15)       # the edge slabs would leave fewer than 2 columns per interior rank
    with pytest.raises(ob.ObmdError):
        ob.slab_columns(7, 4, 0)            # fewer than 2 columns per rank
    rng = np.random.default_rng(1)
    L = 134.37849944008956
    x = rng.uniform(0, L, 20000); x[:3] = (0.0, L, L - 1e-12)
    owner = np.zeros(len(x), int) - 1
    for r in range(4):
        lo, hi = ob.slab_bounds(0.0, L, 1.0, 4, r)
        m = ob.slab_mask(x, lo, hi, r == 3, L)
        assert np.all(owner[m] == -1)
        owner[m] = r
    assert np.all(owner >= 0)


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import obmd_b200 as ob
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
uid = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
assert uid[0] == bytes(range(128))
L = 67.18924972004478
rng = np.random.default_rng(3); x = rng.uniform(0, L, 5000)
lo, hi = ob.slab_bounds(0.0, L, 1.0, world, rank)
m = ob.slab_mask(x, lo, hi, rank == world - 1, L)
cnt = torch.tensor([float(m.sum()), float(x[m].max() if m.any() else 0.0)], dtype=torch.float64)
tot = cnt[:1].clone(); dist.all_reduce(tot)                        # packed sum
mx = cnt[1:].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)    # packed max
assert int(tot.item()) == len(x) and abs(mx.item() - x.max()) < 1e-15
dist.destroy_process_group()
print("OK", rank)
"""


def test_rendezvous_and_reductions_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("OK %d" % r) in o, o[-2000:]


@pytest.mark.gpu
def test_slab_run_matches_single_gpu():
    """tests/mgpu_check.py under torchrun on 2 GPUs (skipped on a 1-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0 and "MGPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
