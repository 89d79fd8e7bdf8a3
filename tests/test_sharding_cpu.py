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


# ---- row tiles: the boundary-exchange protocol (host logic only; the device side is tests/test_modes_gpu.py) ------------------
def _tile_Hs(n=60, Wc=448, Hc=1024, fw=320, fh=180):
    rng = np.random.default_rng(3)
    Hs, y = [], Hc - fh - 6.0
    for t in range(n):
        y += rng.choice([-37.0, -25.0, 31.0, -90.0, 55.0])          # a camera that wanders across the tile boundaries both ways
        y = float(np.clip(y, 4.0, Hc - fh - 6.0))
        T = np.eye(3); T[0, 2] = (Wc - fw) / 2 + 30.0 * np.sin(t / 3.0); T[1, 2] = y
        Hs.append(T)
    return Hs


def test_tile_planner_invariants():
    """every blend sees fresh boundary states, hops only happen when something changed, planners in lock step"""
    fw, fh, Wc, Hc = 320, 180, 448, 1024
    a = sh.TilePlanner(Wc, Hc, 4, fw, fh, 208)
    b = sh.TilePlanner(Wc, Hc, 4, fw, fh, 208)
    n_carry = 0
    for H in _tile_Hs():
        was_down, was_up = list(a.stale_down), list(a.stale_up)
        ops = a.plan(H)
        assert ops == b.plan(H)                                       # deterministic: two processes derive the same list
        blends = [o[1] for o in ops if o[0] == "blend"]
        assert 1 <= len(blends) <= 2
        fresh_d, fresh_u = list(was_down), list(was_up)
        for o in ops:
            if o[0] == "carry":
                _, src, dst, up, block = o
                n_carry += 1
                assert abs(src - dst) == 1 and dst == (src - 1 if up else src + 1)
                assert (was_up if up else was_down)[src]              # only stale states are refreshed
                if up:
                    assert src == 3 or not fresh_u[src + 1]           # the sender's own incoming state was refreshed first
                    fresh_u[src] = False
                else:
                    assert src == 0 or not fresh_d[src - 1]
                    fresh_d[src] = False
                assert 0 <= block < (a.ext[src][1] - a.ext[src][0] + 15) // 16
            elif o[0] == "blend":
                g = o[1]
                assert g == 0 or not fresh_d[g - 1]                   # ghost_top of g is fresh when g blends
                assert g == 3 or not fresh_u[g + 1] or a.ext[g][1] >= Hc
            else:
                _, g, n, x0, ya, w, h = o
                assert g in blends and n not in blends and abs(g - n) == 1
                assert a.own[g][0] <= ya and ya + h <= a.own[g][1]    # only the owner's own rows travel ...
                assert a.ext[n][0] <= ya and ya + h <= a.ext[n][1]    # ... into the neighbour's halo
    assert n_carry >= 10
    with pytest.raises(ValueError):
        sh.TilePlanner(Wc, Hc, 4, fw, fh, 100)                        # not a multiple of 16
    with pytest.raises(ValueError):
        sh.TilePlanner(Wc, Hc, 4, fw, fh, 256)                        # halo reaches beyond the adjacent tile
    with pytest.raises(ValueError):
        sh.TilePlanner(Wc, Hc, 4, fw, fh, 64).plan(_tile_Hs(1)[0] @ np.diag([1.0, 1.0, 1.0]) + np.array([[0, 0, 0], [0, 0, -500.0], [0, 0, 0]]))


def _tile_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pl = sh.TilePlanner(448, 1024, world, 320, 180, 208)
    log = []
    for H in _tile_Hs(40):
        for op in pl.plan(H):
            if op[0] == "blend":
                continue
            src, dst = op[1], op[2]
            payload = torch.tensor([(len(op[0]) * 7919 + sum(int(v) * (i + 1) for i, v in enumerate(op[1:]))) % 1000003], dtype=torch.int64)
            if src == rank:
                dist.send(payload, dst=dst)
                log.append(("s", dst, int(payload)))
            elif dst == rank:
                got = torch.zeros(1, dtype=torch.int64)
                dist.recv(got, src=src)
                assert int(got) == int(payload)                      # the matching operation of the sender's identical plan
                log.append(("r", src, int(got)))
    dist.barrier()
    q.put((rank, len(log)))
    dist.destroy_process_group()


def test_tile_protocol_gloo_world2():
    """two processes, one tile each: every send of the shared plan meets its recv in order (no deadlock, same payload)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    counts = dict(q.get(timeout=5) for _ in range(2))
    assert counts[0] == counts[1] and counts[0] > 0
