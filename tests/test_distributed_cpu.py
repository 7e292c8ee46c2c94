"""N>1 path on CPU: world_size-2 gloo run of the frame-sharding + single-reduction logic
(handheld_super_resolution.distributed), with the NumPy oracle standing in for the per-frame device work.
Checks that sharded accumulation + ONE sum equals the sequential accumulation up to float32 summation order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"), os.path.join(ROOT, "oracle"), HERE):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hhsr_oracle as O
    from helpers import CFA, load
    from handheld_super_resolution.distributed import allreduce_accumulators, shard_frames
    st = load("stage_cases.npz")
    rng = np.random.default_rng(0)
    n_frames = 5
    H, W = st["raw"].shape
    frames = [(st["raw"] * (0.8 + 0.05 * k)).astype(np.float32) for k in range(n_frames)]
    flows = [(st["flow_irreg"] + rng.uniform(-0.5, 0.5, st["flow_irreg"].shape)).astype(np.float32) for _ in range(n_frames)]
    rs = [rng.random((H, W)).astype(np.float32) for _ in range(n_frames)]
    flat = torch.zeros(2 * 2 * H * 2 * W * 3, dtype=torch.float32)          # num || den in one buffer
    num, den = flat[:flat.numel() // 2].view(2 * H, 2 * W, 3), flat[flat.numel() // 2:].view(2 * H, 2 * W, 3)
    acc = torch.zeros((H, W), dtype=torch.float64)
    for k in shard_frames(n_frames, rank, world):
        O.accumulate(frames[k], flows[k], st["covs1"], rs[k], num.numpy(), den.numpy(), CFA, 2, 32)
        acc += torch.from_numpy(rs[k]).double()
    allreduce_accumulators(num, den, acc)
    if rank == 0:
        n1 = np.zeros((2 * H, 2 * W, 3), np.float32)
        d1 = np.zeros_like(n1)
        for k in range(n_frames):
            O.accumulate(frames[k], flows[k], st["covs1"], rs[k], n1, d1, CFA, 2, 32)
        q.put((float(np.abs(num.numpy() - n1).max()), float(np.abs(den.numpy() - d1).max()),
               float(np.abs(acc.numpy() - sum(r.astype(np.float64) for r in rs)).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_accumulation_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    dn, dd, da = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert dn < 1e-5 and dd < 1e-5          # float32 summation order only (1e-5)
    assert da < 1e-12


def test_p2p_row_slices_tile_the_image():
    """Row slices of the fused peer-memory reduction: contiguous, cover [0, Hs), rank 0 about half as tall."""
    from handheld_super_resolution.distributed import p2p_row_slices
    for world in (1, 2, 3, 4, 8, 16):
        for Hs in (257, 6000, 9000, 12288):
            sl = p2p_row_slices(Hs, world)
            assert sl[0][0] == 0 and sl[-1][1] == Hs and len(sl) == world
            assert all(a < b for a, b in sl) and all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
            if world > 1 and Hs >= 6000:
                h0, h1 = sl[0][1] - sl[0][0], sl[1][1] - sl[1][0]
                assert abs(h0 - h1 / 2) <= 1
