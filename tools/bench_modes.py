#!/usr/bin/env python
"""Sharded modes of the stitching path (SURVEY.md 8e / BASELINE.json configs 3-5) -- one JSON line per run.

    python tools/bench_modes.py --mode streams [--streams 64] [--frames 16] [--size 1280x720] [--detector orb]
    python tools/bench_modes.py --mode pairs   [--frames 64] [--size 1920x1080] [--detector sift]
    python tools/bench_modes.py --mode tiles   [--frames 48] [--size 3840x2160] [--canvas 16384x16384]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_modes.py --mode ...

streams : independent streams, stream s -> rank s % N, no data-path collective (config 4).  All streams of a rank are advanced
          together: bm_process_frame_begin on every handle, then bm_process_frame_end on every handle, so their kernels overlap.
pairs   : offline frame-pair sharding (config 3 at N GPUs): contiguous chunks of pairs per rank (1-frame halo), per-pair
          detect / match / RANSAC on the device, ONE all_gather of the 3x3 relative homographies (72 B per pair), then the
          reference's sequential validate / smooth / prefix composition on every rank.
tiles   : canvas row tiles (config 5): rank g owns rows [g*Hc/N, (g+1)*Hc/N) and warps + blends every frame whose window touches
          its tile (homographies known: the sweep's ground truth), tiles gathered with one NCCL all_gather at the end.
Timing: barrier + synchronize on both sides, max over ranks; frames are synthetic (b200mosaic.synth).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import cv2

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class Ctx:
    """process-group context shared by the mode runners (bench.py builds one too)"""

    def __init__(self, rank, world, local, dist, torch):
        self.rank, self.world, self.local, self.dist, self.torch = rank, world, local, dist, torch

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])


def run_streams(ctx, streams=64, frames=16, warmup=3, size="1280x720", det="orb", threads=2, ahead=3):
    """config 4: `streams` concurrent streams, stream s -> rank s % N, no data-path collective"""
    import b200mosaic
    from b200mosaic import sharding as sh
    from b200mosaic.synth import DroneSweep, make_ground
    torch = ctx.torch
    w, h = map(int, size.split("x"))
    n = frames + warmup + 1
    ground = make_ground(4096, 2000)
    nseq = 8                                                 # distinct sweeps; stream s replays sweep s % 8 (content is irrelevant for throughput)
    seqs = [torch.from_numpy(np.stack(DroneSweep(w, h, seed=2000 + q, ground=ground, max_step=10.0, max_travel=0.6 * h).frames(n))).pin_memory()
            for q in range(nseq)]
    fb = h * w * 3
    mine = sh.shard_streams(streams, ctx.rank, ctx.world)
    vms = [b200mosaic.VideMosaic(seqs[s % nseq][0].numpy(), detector_type=det, show_intermediate=False, visualize=False, device=ctx.local) for s in mine]

    for vm in vms:
        vm.warm_up()                                         # setup: every detector graph captured up front (executes nothing)
    # A rank drives its streams from `nthr` host threads (the library releases the GIL; a handle is only ever touched by its own thread):
    # one thread issues ~30 driver calls per frame and saturates at ~7000 frames/s, less than the GPU can do with 720p ORB frames.
    # Every stream also hands over its next `ahead` frames (bm_prefetch_frame), like the single-stream headline does.
    nthr = max(1, min(threads, len(vms)))
    groups = [list(zip(vms, mine))[g::nthr] for g in range(nthr)]

    def step(group, i):
        for vm, s in group:
            base = seqs[s % nseq].data_ptr()
            vm.begin_frame_ptr(base + i * fb)
            for k in range(1, ahead + 1):
                if i + k < n:
                    vm.prefetch_ptr(base + (i + k) * fb)
        return [vm.end_frame() for vm, _ in group]

    def drive(group, lo, hi, out):
        ok = 0
        for i in range(lo, hi):
            ok += sum(1 for st in step(group, i) if st == 0)
        for vm, _ in group:
            vm.sync()
        out.append(ok)

    def run(lo, hi):
        import threading
        out = []
        th = [threading.Thread(target=drive, args=(g, lo, hi, out)) for g in groups[1:]]
        for t in th:
            t.start()
        drive(groups[0], lo, hi, out)
        for t in th:
            t.join()
        return sum(out)
    run(1, warmup + 1)
    ctx.barrier()
    t0 = time.perf_counter()
    ok = run(warmup + 1, n)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    for vm in vms:
        vm.close()
    total = streams * frames
    return {"mode": "streams", "scaling": "strong", "metric": f"aggregate mosaic frames/sec over {streams} concurrent {w}x{h} {det.upper()} streams",
            "value": total / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt / frames, "steps": frames, "warmup": warmup,
            "config": {"workload": f"{streams} streams x {frames} frames, {w}x{h}, detector={det}, stream s -> rank s % {ctx.world}, "
                                   f"{len(mine)} streams on rank 0 driven by {nthr} host thread(s), begin/end interleaved, {ahead} frames staged ahead per stream, "
                                   f"host frames in pinned memory (H2D inside)",
                       "frames_ok_rank0": ok, "collective": "none"}}


def run_pairs(ctx, frames=64, size="1920x1080", det="sift"):
    """config 3 offline at N GPUs: contiguous chunks of frame pairs per rank, ONE all_gather of the relative homographies"""
    import b200mosaic
    from b200mosaic import sharding as sh
    from b200mosaic.synth import DroneSweep
    torch = ctx.torch
    w, h = map(int, size.split("x"))
    n = frames + 1
    pinned = torch.from_numpy(np.stack(DroneSweep(w, h, seed=1234, ground_size=4096, max_step=12.0, max_travel=0.8 * h).frames(n))).pin_memory()
    fr = [pinned[t].numpy() for t in range(n)]               # views into pinned memory: DMA'd directly
    s, e = sh.shard_pairs(n, ctx.rank, ctx.world)
    # handle creation (allocations, graph capture) and the chunk's first pair are the untimed warm-up
    vm = b200mosaic.VideMosaic(fr[s - 1], detector_type=det, show_intermediate=False, visualize=False, device=ctx.local) if e > s else None
    if vm is not None:
        vm.warm_up()                                         # every detector graph captured now (executes nothing), not inside the timed pairs
    st0, Hs0 = sh.estimate_pairs(fr, s, min(s + 1, e), detector_type=det, device=ctx.local, vm=vm)
    sh.all_gather_pairs(sh.pack_pairs(st0, Hs0), n, ctx.rank, ctx.world, ctx.dist, device="cuda")      # untimed: first use of the collective
    ctx.barrier()
    t0 = time.perf_counter()
    st, Hs = sh.estimate_pairs(fr, s + 1, e, detector_type=det, device=ctx.local, vm=vm)
    st, Hs = st0 + st, Hs0 + Hs
    torch.cuda.synchronize()
    t_est = ctx.max_over_ranks(time.perf_counter() - t0)
    rows = sh.all_gather_pairs(sh.pack_pairs(st, Hs), n, ctx.rank, ctx.world, ctx.dist, device="cuda")
    rel = sh.unpack_pairs(rows)
    H0 = np.eye(3); H0[0, 2] = int(1.2 * w) / 2 - w / 2; H0[1, 2] = int(2 * h) - h
    Habs = sh.compose_chain(H0, rel)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    if vm is not None:
        vm.close()
    timed = n - 1 - ctx.world
    return {"mode": "pairs", "scaling": "strong",
            "metric": f"frame pairs/sec (detect + match + RANSAC, {det.upper()}, {w}x{h}) sharded over ranks + all_gather + prefix composition",
            "value": timed / dt, "unit": "pairs/s", "ms_per_step": 1e3 * dt / timed, "steps": timed, "warmup": 1,
            "config": {"workload": f"{n} frames, pairs [{s},{e}) on rank 0 of {ctx.world}", "pairs_ok": int(sum(1 for r in rel if r is not None)),
                       "composed": int(sum(1 for H in Habs if H is not None)), "collective": "all_gather of 80 B per pair (NCCL)",
                       "all_gather_plus_compose_ms": 1e3 * (dt - t_est)}}


def run_tiles(ctx, frames=48, warmup=3, size="3840x2160", canvas="16384x16384"):
    """config 5: canvas row tiles, every rank warps + blends the frames that touch its rows (boundary rows exchanged with the
    neighbours, see sharding.TileExchange), final NCCL all_gather of the tiles"""
    import b200mosaic
    from b200mosaic import sharding as sh
    from b200mosaic.synth import DroneSweep, make_ground
    torch = ctx.torch
    w, h = map(int, size.split("x"))
    Wc, Hc = map(int, canvas.split("x"))
    n = frames + warmup + 1
    ground = make_ground(4096, 77)
    nseq = min(n, 12)                                        # frame CONTENT is recycled (irrelevant for throughput); poses are not
    base = DroneSweep(w, h, seed=77, ground=cv2.resize(ground, (2 * 4096, 2 * 4096)) if max(w, h) > 3000 else ground, max_step=40.0,
                      noise_sigma=2.0).frames(nseq)
    pinned = torch.from_numpy(np.stack(base)).pin_memory()   # host frames in pinned memory (DMA'd directly, like the other modes)
    fr = [pinned[t % nseq].numpy() for t in range(n)]
    y0, y1 = sh.tile_rows(Hc, ctx.rank, ctx.world)
    # the camera climbs the whole canvas in n frames (so every row tile gets work), with a slow sideways weave and rotation
    Hs = []
    for t in range(n):
        ang = np.deg2rad(2.0 * np.sin(t / 7.0))
        R = np.array([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]])
        T = np.eye(3); T[0, 2] = Wc / 2 - w / 2 + 0.05 * Wc * np.sin(t / 5.0); T[1, 2] = (Hc - h - 8) * (1.0 - t / max(n - 1, 1)) + 4
        Hs.append(T @ R)
    # halo: the tallest warped frame (+-2 degrees of rotation) + slack, a multiple of 16 rows
    halo = -(-int(h * np.cos(np.deg2rad(2.0)) + w * np.sin(np.deg2rad(2.0)) + 40) // 16) * 16
    tiler = sh.TileGroup(fr[0], Wc, Hc, ctx.world, [ctx.rank], halo, ctx.dist, device=ctx.local)
    for t in range(0, warmup + 1):
        tiler.put(fr[t], Hs[t])
    tiler.sync()
    full = torch.zeros((Hc, Wc, 3), dtype=torch.uint8, device=f"cuda:{ctx.local}")       # setup: the gather's destination, allocated and touched up front
    _ = tiler.tile_tensor(); del _                           # setup: the export buffer of the tile comes out of torch's caching allocator afterwards
    torch.cuda.synchronize()
    ctx.barrier()
    t0 = time.perf_counter()
    mine = sum(tiler.put(fr[t], Hs[t]) for t in range(warmup + 1, n))
    tile = tiler.tile_tensor()
    torch.cuda.synchronize()
    t_warp = ctx.max_over_ranks(time.perf_counter() - t0)
    full = sh.gather_tiles(tile, Hc, ctx.rank, ctx.world, ctx.dist, out=full)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    nbytes = int(full.numel())
    hops, rect_bytes = tiler.hops, tiler.rect_bytes
    del full, tile
    tiler.close()
    return {"mode": "tiles", "scaling": "strong",
            "metric": f"frames/sec warped + blended into a {Wc}x{Hc} canvas sharded in {ctx.world} row tiles, incl. the final NCCL all_gather",
            "value": frames / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt / frames, "steps": frames, "warmup": warmup,
            "config": {"workload": f"{w}x{h} frames, tile rows [{y0},{y1}) on rank 0, {mine} of {frames} frames touch it",
                       "warp_blend_fps": frames / t_warp, "gather_ms": 1e3 * (dt - t_warp), "canvas_bytes": nbytes,
                       "collective": "per-frame neighbour exchange of boundary rows (NCCL send/recv) + final all_gather of the tiles",
                       "boundary_exchange": True, "halo_rows": halo, "carry_hops": hops, "halo_rect_bytes": rect_bytes}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["streams", "pairs", "tiles"])
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--frames", type=int, default=16, help="timed frames (per stream / in total)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", default=None)
    ap.add_argument("--canvas", default="16384x16384")
    ap.add_argument("--detector", default=None)
    ap.add_argument("--threads", type=int, default=2, help="streams mode: host threads per rank")
    ap.add_argument("--ahead", type=int, default=3, help="streams mode: frames staged ahead per stream")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Ctx(rank, world, local, dist, torch)
    if args.mode == "streams":
        out = run_streams(ctx, args.streams, args.frames, args.warmup, args.size or "1280x720", args.detector or "orb", args.threads, args.ahead)
    elif args.mode == "pairs":
        out = run_pairs(ctx, args.frames, args.size or "1920x1080", args.detector or "sift")
    else:
        out = run_tiles(ctx, args.frames, args.warmup, args.size or "3840x2160", args.canvas)
    out.update({"n_gpus": world, "data": "synthetic", "higher_is_better": True})
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
