"""Host-side logic of the multi-GPU modes on CPU: partitioning, the sequential compose chain against the
oracle's control flow, and the two collectives (all_gather of per-pair homographies, gather of canvas tiles) with
world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200mosaic  # noqa: F401
from b200mosaic import sharding as sh
from oracle.mosaic_ref import RefMosaic


def test_partitions_cover_everything():
    for world in (1, 2, 3, 8):
        assert sorted(sum((sh.shard_streams(64, r, world) for r in range(world)), [])) == list(range(64))
        for n in (1, 2, 7, 2000):
            spans = [sh.shard_pairs(n, r, world) for r in range(world)]
            got = sum((list(range(s, e)) for s, e in spans), [])
            assert got == list(range(1, n))
        for hc in (2160, 32768, 100):
            rows = [sh.tile_rows(hc, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == hc
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
            assert all(y0 % 16 == 0 for y0, y1 in rows if y1 > y0)


def test_compose_chain_equals_reference_control_flow():
    rng = np.random.default_rng(0)
    frame = np.zeros((64, 96, 3), np.uint8); frame[8:40, 8:60] = 200
    ref = RefMosaic(frame, detector_type="orb")
    H0 = ref.H_old.copy()
    rel, want = [], []
    for t in range(12):
        H = np.eye(3)
        H[:2, :2] += rng.normal(size=(2, 2)) * 0.01
        H[0, 2], H[1, 2] = rng.normal() * 6, -8 + rng.normal() * 3
        if t == 4:
            H[0, 2] = 80.0            # rejected: translation > 50 -> identity substituted
        if t == 7:
            rel.append(None); want.append(None); continue      # skipped pair: state not advanced
        rel.append(H)
        Hv = H if ref.validate_homography(H) else np.eye(3)
        ref.H_old = ref.H_old @ ref.smooth_homography(Hv)
        want.append(ref.H_old.copy())
    got = sh.compose_chain(H0, rel)
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b)


def test_tile_homography_and_window():
    H = np.array([[1, 0, 100.0], [0, 1, 5000.0], [0, 0, 1]])
    assert sh.touches_tile(H, 1920, 1080, 4096, 8192)
    assert not sh.touches_tile(H, 1920, 1080, 8192, 12288)
    Ht = sh.tile_homography(H, 4096)
    assert Ht[1, 2] == 904.0 and Ht[0, 2] == 100.0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # --- per-pair all_gather: every rank "estimates" its chunk (synthetic H = f(t)) ---
    s, e = sh.shard_pairs(n_frames, rank, world)
    st = [sh.OK if t % 5 else sh.SKIP_FEW for t in range(s, e)]
    Hs = [np.eye(3) * (1 + t) if t % 5 else None for t in range(s, e)]
    rows = sh.all_gather_pairs(sh.pack_pairs(st, Hs), n_frames, rank, world, dist)
    # --- canvas tiles gather ---
    hc, wc = 100, 8
    y0, y1 = sh.tile_rows(hc, rank, world)
    tile = torch.full((y1 - y0, wc, 3), rank + 1, dtype=torch.uint8)
    tile[:, 0, 0] = torch.arange(y0, y1, dtype=torch.uint8)
    full = sh.gather_tiles(tile, hc, rank, world, dist)
    q.put((rank, rows, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_collectives():
    world, n_frames = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, full in res:
        assert rows.shape == (n_frames - 1, 10)
        rel = sh.unpack_pairs(rows)
        for t in range(1, n_frames):
            if t % 5:
                assert np.array_equal(rel[t - 1], np.eye(3) * (1 + t))
            else:
                assert rel[t - 1] is None
        assert full.shape == (100, 8, 3)
        assert np.array_equal(full[:, 0, 0], np.arange(100, dtype=np.uint8))
        y0, y1 = sh.tile_rows(100, 1, 2)
        assert (full[:y0, 1:, :] == 1).all() and (full[y0:, 1:, :] == 2).all()
