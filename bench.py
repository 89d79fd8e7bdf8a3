#!/usr/bin/env python
"""bench.py -- mosaic frames/sec on the synthetic 1080p drone sweep (BASELINE.json configs[2]) + warp/blend roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--detector sift|orb] [--size WxH]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `process_frame` of the stitching hot path on the next synthetic frame of the sweep.
 * value      : frames/s with the frames already resident in HBM (packed BGR), every stage on the device, CUDA-event timed.
 * e2e        : the same through the reference-facing call (bm_process_frame on HOST frames in pinned memory: H2D inside the
                timed region, the per-frame result (status, H) read back every step, the canvas fetched once at the end).
 * roofline   : the warp/blend chain (graded kernel group): algorithmic bytes 3N + 6A per frame over its
                CUDA-event time on the launching stream, against MEASURED_PEAKS.json's HBM copy bandwidth.
 * cpu_baseline: the oracle's cv2 path (same calls as the reference's VideMosaic) on the first frames, all host threads.
 * regions    : every leg times R regions of exactly K steps (barrier + synchronize on both sides, max over ranks per region);
                value / e2e are the MEDIAN region, min / max and the per-rank times are reported next to them.
 * <other detector>: the same legs for the detector that is not the headline (the metric names SIFT & ORB).
 * roofline_pyramid: the SIFT Gaussian + DoG pyramid alone (256 N algorithmic bytes per frame).
 * long_run   : configs[2] at its named length (2000 frames) end to end, one region on rank 0.
 * clip       : BASELINE configs 1 / 2 -- the reference's own 592-frame clip through main.main()'s loop with the launcher's swaps, H.264
                decode, finalisation and mosaic.jpg inside the wall-clock time; decode alone and the CPU port beside it (tools/bench_clip.py).
 * modes      : the sharded modes of SURVEY 8e over the N ranks of this launch -- 64 x 720p ORB streams (config 4), offline frame-pair
                sharding with its all_gather (config 3 at N GPUs), 4K frames into a 32768^2 canvas in N row tiles + NCCL gather (config 5).
N > 1: one process per GPU, each rank stitches its own independent sweep (streams sharded one per GPU);
no data-path collective; weak scaling; time = max over ranks.
`--impl reference` times the reference's CPU path (oracle.mosaic_ref.RefMosaic: the same cv2/NumPy calls as main.py) on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
METRIC = "mosaic frames/sec at 1080p (SIFT & ORB); warp/blend HBM GB/s vs peak"      # BASELINE.json `metric`; value = frames/s of --detector


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--detector", default=os.environ.get("BM_BENCH_DETECTOR", "sift"), choices=["sift", "orb"])
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--cpu-frames", type=int, default=12, help="frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--regions", type=int, default=0, help="timed regions of exactly --steps steps each (0: min(5, 360 // steps))")
    ap.add_argument("--single-detector", action="store_true", help="skip the sub-record of the other detector")
    ap.add_argument("--no-modes", action="store_true", help="skip the sharded-mode records (configs 3 offline / 4 / 5)")
    ap.add_argument("--no-long-run", action="store_true", help="skip the 2000-frame run of configs[2] at its named length")
    ap.add_argument("--no-clip", action="store_true", help="skip the real-clip record (configs 1 / 2 end to end with decode)")
    return ap.parse_args()


def make_frames(w, h, n, seed):
    from b200mosaic.synth import DroneSweep
    sweep = DroneSweep(w, h, seed=seed, ground_size=max(4096, 2 * max(w, h)), max_step=12.0, max_travel=0.8 * h)
    return sweep.frames(n), sweep


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


def cpu_reference_fps(frames, detector, nthreads, stages=None):
    """The reference's CPU path (cv2 + NumPy, same calls as main.py's VideMosaic) on the given frames; frames/s.
    `stages` (a dict) receives the per-stage split SURVEY.md 8d asks for: seconds inside detectAndCompute, match, findHomography and
    warp() (warpPerspective + blend) summed over the frames -- measured by timing wrappers around the port's own calls."""
    import cv2
    from oracle.mosaic_ref import RefMosaic
    cv2.setNumThreads(nthreads)
    cv2.ipp.setUseIPP(True)                       # timing runs keep IPP on; parity runs switch it off
    m = RefMosaic(frames[0], detector_type=detector)
    if stages is not None:
        acc = {"detectAndCompute": 0.0, "match": 0.0, "findHomography": 0.0, "warp": 0.0}

        def timed(name, fn):
            def w(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    acc[name] += time.perf_counter() - t
            return w

        class Det:                                # cv2 feature objects do not take attributes: a pass-through with one timed method
            def __init__(self, d):
                self.detectAndCompute = timed("detectAndCompute", d.detectAndCompute)
        m.detector = Det(m.detector)
        m.match = timed("match", m.match)
        m.findHomography = timed("findHomography", m.findHomography)
        m.warp = timed("warp", m.warp)
    t0 = time.perf_counter()
    for i, f in enumerate(frames[1:], 1):
        m.process_frame(f, i)
    dt = time.perf_counter() - t0
    if stages is not None:
        nfr = max(len(frames) - 1, 1)
        stages.update({k: 1e3 * v / nfr for k, v in acc.items()})
        stages["other"] = 1e3 * (dt - sum(acc.values())) / nfr
        stages["unit"] = "ms per frame"
    return (len(frames) - 1) / dt, dt


def run_reference(args, rank, world):
    w, h = map(int, args.size.lower().split("x"))
    if rank != 0:
        return
    n = args.steps + args.warmup + 1
    frames, _ = make_frames(w, h, n, 1234)
    import cv2
    from oracle.mosaic_ref import RefMosaic
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    import contextlib
    import io
    m = RefMosaic(frames[0], detector_type=args.detector)
    with contextlib.redirect_stdout(io.StringIO()):          # the reference's warnings (main.py:723-798) must not break the one-line contract
        for i in range(1, args.warmup + 1):
            m.process_frame(frames[i], i)
        t0 = time.perf_counter()
        for i in range(args.warmup + 1, n):
            m.process_frame(frames[i], i)
        dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/f64 (cv2 CPU)", "data": "synthetic",
            "config": workload_config(args, w, h, frames[0]),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} frames of the same sweep, oracle.mosaic_ref.RefMosaic (cv2 {cv2.__version__}, "
                                       f"the reference's VideMosaic calls minus its display-only canvas copy / draw_border / gc.collect, so the port is "
                                       f"slightly FASTER than the reference), cv2.setNumThreads({cores}), IPP on"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, w, h, frame0):
    ch, cw = int(2 * h), int(1.2 * w)
    return {"workload": f"synthetic {w}x{h} drone sweep (seed 1234, <=12 px/frame drift), detector={args.detector}, "
                        f"canvas {cw}x{ch} (reference defaults 2x / 1.2x), one process_frame per step",
            "detector": args.detector, "frame": [h, w], "canvas": [ch, cw],
            "l2": ("inputs larger than L2: every step streams the frame's 0.49 GB Gaussian + DoG pyramid (SIFT) through the 126 MB L2, "
                   "no explicit flush" if args.detector == "sift" else
                   "every step processes a new 6.2 MB frame, its 8-level pyramid and a different canvas window (~60 MB working set "
                   "< 126 MB L2); inputs differ every step, no explicit L2 flush -- the ORB path is L2 resident by design")}


def pin_rank_cores(local_rank, world):
    """each rank's launch thread on its own cores: the per-frame loop is host driven, and 8 ranks x (python + driver threads)
    otherwise migrate over the box's cores (8-GPU efficiency 0.88 with max-over-ranks on a 13 ms window)"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return cores, mine
    except (AttributeError, OSError):
        return None, None


def ramp_clocks(torch, ms=120.0):
    """untimed: keep the SMs busy for ~ms so that the timed region does not start on an idle-clocked GPU (host-side setup such as
    page-locking hundreds of MB leaves the device idle for 100s of ms; the N=1 e2e leg of round 1 measured 16 % low for that reason)"""
    a = torch.empty((4096, 4096), device="cuda", dtype=torch.bfloat16).normal_()
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < ms:
        for _ in range(8):
            a = (a @ a).clamp_(-1, 1)
        torch.cuda.synchronize()


def regions_max_over_ranks(torch, dist, times):
    """times: this rank's per-region seconds.  Returns (per-region max over ranks, all ranks' rows)."""
    t = torch.tensor(times, device="cuda", dtype=torch.float64)
    if dist is None:
        return list(times), [list(times)]
    rows = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(rows, t)
    allr = torch.stack(rows)
    return [float(v) for v in allr.max(dim=0).values], [[float(v) for v in r] for r in allr]


LOOKAHEAD = int(os.environ.get("BM_AHEAD", "3"))      # frames handed over ahead of the current one (the library stages up to 3)


def ahead_ptrs(base, i, fb, n):
    """pointers of the frames i+1 .. i+LOOKAHEAD of a contiguous frame array (None past its end)"""
    return [base + (i + k) * fb if (k <= LOOKAHEAD and i + k < n) else None for k in (1, 2, 3)]


def headline(args, det, frames, dev_frames, pinned, ctx, lib, want_chain, sample_clocks):
    """legs 1 (frames resident in HBM), 1b (chain alone, roofline) and 2 (end to end from pinned host frames) for one detector;
    every leg is R regions of exactly K steps, each region bracketed by barrier + synchronize, timed per rank, max over ranks"""
    import b200mosaic
    torch, dist, rank, local_rank = ctx.torch, ctx.dist, ctx.rank, ctx.local
    h, w = frames[0].shape[:2]
    fb = h * w * 3
    K, W, R = args.steps, args.warmup, args.regions
    n = len(frames)
    out = {}
    # ---------------- leg 1: frames resident in HBM ----------------
    vm = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False, device=local_rank)
    base = dev_frames.data_ptr()
    vm.warm_up()                                   # setup: CUDA graphs of the detector captured up front (executes nothing)
    ramp_clocks(torch)
    for i in range(1, W + 1):
        vm.process_frame_device(base + i * fb, *ahead_ptrs(base, i, fb, n))
    vm.sync()
    statuses, t_dev, ev_ms = [], [], []
    launches0 = lib.bm_kernel_launches()
    sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
    i = W + 1
    for r in range(R):
        ctx.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(K):
            statuses.append(vm.process_frame_device(base + i * fb, *ahead_ptrs(base, i, fb, n)))
            i += 1
        vm.sync()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ev_ms.append(e0.elapsed_time(e1))
        t_dev.append(max(wall, ev_ms[-1] * 1e-3))         # the step loop is host-driven; wall >= device span
    out["launches"] = (lib.bm_kernel_launches() - launches0) / R
    out["clocks"] = sampler.stop() if sampler else None
    out["frames_ok"] = sum(1 for s in statuses if s == 0)
    out["t_dev"], out["ev_ms"] = t_dev, ev_ms
    del vm

    # ---------------- leg 1b: the warp/blend chain alone (roofline) ----------------
    # same frames and pipeline, but detect of frame t+1 is ordered after the chain of frame t (no overlap), so the CUDA events
    # around the chain on its launching stream measure the chain and nothing else
    # (rank 0 only, the others wait at the barrier below: with 8 ranks launching at once on 4 host cores each, the host could not keep
    # the idle GPU fed and the event interval measured launch gaps -- 0.32 ms instead of 0.126)
    if want_chain and rank == 0:
        vmr = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False, device=local_rank)
        vmr.set_overlap(False)
        Kr = min(n - W - 2, 60)
        for i in range(1, W + 1):
            vmr.process_frame_device(base + i * fb)
        vmr.sync()
        vmr.timing(enable=True, reset=True)
        for i in range(W + 1, W + 1 + Kr):
            vmr.process_frame_device(base + i * fb)
        vmr.sync()
        out["chain"] = vmr.timing(reset=True)
        del vmr
    if want_chain:
        ctx.barrier()

    # ---------------- leg 2: end to end through the host-facing call ----------------
    vm2 = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False, device=local_rank)
    pbase = pinned.data_ptr()
    canvas_host = torch.empty(tuple(vm2.output_img.shape), dtype=torch.uint8).pin_memory().numpy()     # setup: reusable host buffer
    vm2.warm_up()
    ramp_clocks(torch)
    for i in range(1, W + 1):
        vm2.process_frame_ptr(pbase + i * fb, *ahead_ptrs(pbase, i, fb, n))
    vm2.sync()
    t_e2e = []
    i = W + 1
    canvas = None
    for r in range(R):
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            # H2D (the copies of frames i+1 .. i+3 are started while frame i is processed: what a reader thread three frames ahead of
            # the stitcher provides, run.AheadCapture) + all kernels + D2H of (counts, H)
            vm2.process_frame_ptr(pbase + i * fb, *ahead_ptrs(pbase, i, fb, n))
            i += 1
        canvas = vm2.read_canvas(canvas_host)          # the canvas (what becomes mosaic.jpg) D2H into the caller's pinned buffer, every region
        torch.cuda.synchronize()
        t_e2e.append(time.perf_counter() - t0)
    out["t_e2e"] = t_e2e
    out["canvas"] = canvas
    out["vm2"] = vm2
    return out


def long_run(det, frames, pinned, ctx, total=2000):
    """BASELINE configs[2] at its NAMED length: 2000 frames through the host-facing call (pinned host frames, H2D + per-frame read-back
    inside the timed region), one region, rank 0.  The frames are the sweep's resident ones walked forwards and backwards (a camera that
    flies the same serpentine to and fro: consecutive frames stay <= 24 px apart, so every frame validates); a turn skips one frame so
    that no frame appears twice among the current frame and the two staged right behind it (a third-ahead frame that equals the current
    one is simply staged one call later)."""
    import b200mosaic
    torch = ctx.torch
    h, w = frames[0].shape[:2]
    fb = h * w * 3
    n = len(frames)
    seq, i, d = [], 1, 1
    while len(seq) < total + 3:
        seq.append(i)
        if not 0 <= i + d < n:
            d = -d
            i += 2 * d
        else:
            i += d
    vm = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False, device=ctx.local)
    pbase = pinned.data_ptr()
    vm.warm_up()
    ramp_clocks(torch)
    ok = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(total):
        ptrs = [pbase + seq[t + k] * fb if k <= LOOKAHEAD else None for k in (1, 2, 3)]
        ok += vm.process_frame_ptr(pbase + seq[t] * fb, *ptrs) == 0
    vm.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    vm.close()
    return {"what": f"configs[2] at its named length: {total} frames, detector={det}, end to end from pinned host frames (the resident sweep walked "
                    f"to and fro), one timed region on rank 0", "frames": total, "frames_ok": int(ok), "seconds": dt, "e2e_frames_per_s": total / dt}


def stats(xs):
    xs = sorted(xs)
    return {"min": xs[0], "median": xs[len(xs) // 2] if len(xs) % 2 else 0.5 * (xs[len(xs) // 2 - 1] + xs[len(xs) // 2]), "max": xs[-1]}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import contextlib
    import io
    import torch
    import b200mosaic
    from b200mosaic import _lib
    sys.path.insert(0, str(ROOT / "tools"))
    import bench_modes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    all_cores, my_cores = pin_rank_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = bench_modes.Ctx(rank, world, local_rank, dist, torch)
    lib = _lib.load()
    w, h = map(int, args.size.lower().split("x"))
    K, W = args.steps, args.warmup
    if args.regions <= 0:
        args.regions = max(1, min(5, 360 // max(K, 1)))
    R = args.regions
    n = R * K + W + 4
    frames, sweep = make_frames(w, h, n, 1234 + 1000 * rank)
    fb = h * w * 3
    dev_frames = torch.from_numpy(np.stack(frames)).cuda(local_rank)          # (n, h, w, 3) u8
    pinned = torch.from_numpy(np.stack(frames)).pin_memory()

    dets = [args.detector] + ([d for d in ("sift", "orb") if d != args.detector] if not args.single_detector else [])
    res = {}
    for di, det in enumerate(dets):
        res[det] = headline(args, det, frames, dev_frames, pinned, ctx, lib, want_chain=(di == 0), sample_clocks=(di == 0))
    main_det = dets[0]
    hm = res[main_det]
    longrec = None
    if not args.no_long_run:
        if rank == 0:
            try:
                longrec = long_run(main_det, frames, pinned, ctx)
            except Exception as ex:                     # reported, never silently dropped
                longrec = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        ctx.barrier()

    # SURVEY 8f rank 1 / 3 (outside the timed regions): finalisation and preview thumbnail of the final canvas on the device
    vm2 = hm["vm2"]
    vm2.finalize()
    tf = time.perf_counter()
    final_img = vm2.finalize()
    finalize_ms = 1e3 * (time.perf_counter() - tf)
    vm2.finalize_jpeg()
    tf = time.perf_counter()
    jpeg_bytes = vm2.finalize_jpeg()
    finalize_jpeg_ms = 1e3 * (time.perf_counter() - tf)
    vm2.preview()
    tf = time.perf_counter()
    thumb = vm2.preview()
    preview_ms = 1e3 * (time.perf_counter() - tf)
    canvas = hm["canvas"]
    for det in dets:
        res[det].pop("vm2").close()

    # SIFT pyramid alone (roofline_pyramid): its kernels back to back on one stream between two CUDA events
    pyr = mt = None
    if rank == 0:
        import ctypes as C
        gray = torch.from_numpy(np.ascontiguousarray(frames[1][:, :, 1])).cuda(local_rank)
        ms = C.c_double(0); by = C.c_double(0)
        torch.cuda.synchronize()
        if lib.bm_sift_pyramid_ms(C.c_void_p(gray.data_ptr()), h, w, 20, C.byref(ms), C.byref(by)) == 0:
            pyr = (ms.value, by.value)
        mt = None
        if lib.bm_match_l2_ms(700, 700, 50, C.byref(ms), C.byref(by)) == 0:
            mt = (ms.value, by.value)
    del dev_frames

    # ---------------- max over ranks, per region ----------------
    agg = {}
    for det in dets:
        dmax, drows = regions_max_over_ranks(torch, dist, res[det]["t_dev"])
        emax, erows = regions_max_over_ranks(torch, dist, res[det]["t_e2e"])
        agg[det] = {"dev": dmax, "e2e": emax, "dev_rows": drows, "e2e_rows": erows}
    total_frames = K * world

    # ---------------- sharded modes (configs 3 offline / 4 / 5), bounded ----------------
    modes = None
    if not args.no_modes:
        modes = []
        for fn, kw in ((bench_modes.run_streams, dict(streams=64, frames=16, warmup=3)),
                       (bench_modes.run_pairs, dict(frames=max(40, 16 * world) + world)),
                       (bench_modes.run_tiles, dict(frames=24, warmup=2, size="3840x2160", canvas="32768x32768"))):
            try:
                torch.cuda.empty_cache()
                m = fn(ctx, **kw)
                m["n_gpus"] = world
            except Exception as ex:                     # a mode that cannot run is reported, never silently dropped
                m = {"mode": fn.__name__, "error": f"{type(ex).__name__}: {ex}"[:300]}
            modes.append(m)

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        wb_ms, wb_bytes, wb_frames = hm["chain"]
        achieved = (wb_bytes / 1e9) / (wb_ms / 1e3) if wb_ms > 0 else 0.0
        # DRAM traffic of the chain per frame from the committed ncu captures (profiles/): cold-cache `--set full` replay and the
        # in-pipeline figure (range replay without cache control), both per frame like `achieved`
        traffic, traffic_src, traffic_in_pipeline = None, None, None
        for name in ("r02_chain_ncu.json", "r01_chain_ncu.json"):
            tj = ROOT / "profiles" / name
            if tj.exists() and w == 1920 and h == 1080:
                try:
                    j = json.loads(tj.read_text())
                    kk = j["kernels"]
                    per_frame = j.get("launches_per_frame") or {"k_warp_rows": 1, "k_dt_local": 2, "k_dt_diag_chain16": 1, "k_dt_vert_local": 1,
                                                                "k_dt_vert_chain16": 1, "k_dt_weights": 1, "k_blur_blend": 1, "k_rowscan_bgrx": 1}
                    traffic = float(sum(c * (kk[k]["dram_read"] + kk[k]["dram_write"]) for k, c in per_frame.items() if k in kk))
                    traffic_src = f"profiles/{name} (ncu --set full, caches flushed per kernel replay)"
                    traffic_in_pipeline = j.get("in_pipeline_dram_bytes_per_frame")
                    break
                except (KeyError, ValueError):
                    traffic = None
        if all_cores:
            try:
                os.sched_setaffinity(0, all_cores)     # the CPU baseline may use every host core
            except OSError:
                pass
        # configs 1 / 2: the reference's own clip through main.main()'s loop with decode inside (tools/bench_clip.py), rank 0 only; before the CPU
        # baselines below, whose 16 cv2 worker threads keep spinning for a while and would compete with the decode and launch threads
        clip = None
        if not args.no_clip:
            try:
                import bench_clip
                clip = bench_clip.clip_record(cpu_frames=0 if args.no_cpu_baseline else 30, repeat=3)
            except Exception as ex:                     # reported, never silently dropped
                clip = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        cpus = {}
        if not args.no_cpu_baseline:
            import cv2
            cores = os.cpu_count() or 1
            nf = min(args.cpu_frames + 1, len(frames))
            for det in dets:
                st = {}
                with contextlib.redirect_stdout(io.StringIO()):
                    fps, dt = cpu_reference_fps(frames[:nf], det, cores, st)
                cpus[det] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "stages": st,
                             "sample": f"first {nf - 1} frames of the same sweep ({dt:.1f} s), detector={det}, oracle.mosaic_ref.RefMosaic = the "
                                       f"reference's cv2 {cv2.__version__}/NumPy calls minus its display-only copies, cv2.setNumThreads({cores}), IPP on"}
        finalize_cpu_ms = preview_cpu_ms = jpeg_cpu_ms = jpeg_same = None
        if cpus:                                          # the reference's own functions on the same canvas, same host
            from oracle import finalize as ofin
            tc = time.perf_counter()
            ofin.scale_to_screen(ofin.crop_black_areas(canvas, threshold=80, margin=30))
            finalize_cpu_ms = 1e3 * (time.perf_counter() - tc)
            tc = time.perf_counter()
            enc = cv2.imencode(".jpg", final_img)[1]          # what cv2.imwrite('mosaic.jpg', scaled) encodes (main.py:1664-1665)
            jpeg_cpu_ms = 1e3 * (time.perf_counter() - tc)
            jpeg_same = bool(enc.tobytes() == jpeg_bytes)
            try:                                          # gui.py:143-158 on the copy main.py:1630-1632 hands over (host side only)
                from oracle import preview as opv
                tc = time.perf_counter()
                opv.gui_thumbnail(canvas.copy())
                preview_cpu_ms = 1e3 * (time.perf_counter() - tc)
            except ImportError:
                pass
        info_bytes = 160 + 16

        def record(det):
            a = agg[det]
            sd, se = stats(a["dev"]), stats(a["e2e"])
            return {"value": total_frames / sd["median"], "ms_per_step": 1e3 * sd["median"] / K,
                    "e2e": {"value": total_frames / se["median"], "unit": "frames/s", "h2d_bytes_per_step": fb,
                            "d2h_bytes_per_step": info_bytes + int(canvas.nbytes / K)},
                    "regions": {"count": R, "steps_each": K,
                                "value_min_median_max": [total_frames / sd["max"], total_frames / sd["median"], total_frames / sd["min"]],
                                "e2e_min_median_max": [total_frames / se["max"], total_frames / se["median"], total_frames / se["min"]],
                                "per_rank_ms_per_step": [[1e3 * t / K for t in row] for row in a["dev_rows"]],
                                "per_rank_e2e_ms_per_step": [[1e3 * t / K for t in row] for row in a["e2e_rows"]]},
                    "gpu_launches": int(round(res[det]["launches"])), "frames_ok": res[det]["frames_ok"],
                    "event_ms_per_step": float(np.median(res[det]["ev_ms"])) / K, "cpu_baseline": cpus.get(det)}
        rm = record(main_det)
        line = {"metric": METRIC, "value": rm["value"], "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": rm["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 / u32 fixed point (warp, DT, ORB) + f32 (blend weights, SIFT pyramid) + bf16 x bf16 -> f32 tensor cores (SIFT matching, exact) + f64 (RANSAC/LM)", "data": "synthetic",
                "config": workload_config(args, w, h, frames[0]),
                "e2e": rm["e2e"], "gpu_launches": rm["gpu_launches"], "clocks": hm["clocks"],
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                             "traffic_in_pipeline": traffic_in_pipeline,
                             "algorithmic_bytes_per_frame": wb_bytes / max(wb_frames, 1),
                             "kernel": "warp/blend chain (k_warp_rows, k_dt_*, k_blur_blend, k_rowscan_bgrx), 3N+6A bytes per frame; timed with "
                                       "CUDA events on its launching stream in a separate pass without detect overlap",
                             "frames": int(wb_frames),
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                             "ms_per_frame": wb_ms / max(wb_frames, 1)},
                "roofline_pyramid": None if pyr is None else {
                    "bound": "hbm", "achieved": (pyr[1] / 1e9) / (pyr[0] / 1e3), "peak": peak, "unit": "GB/s",
                    "frac": (pyr[1] / 1e9) / (pyr[0] / 1e3) / peak, "algorithmic_bytes_per_frame": pyr[1], "ms_per_frame": pyr[0],
                    "kernel": "SIFT Gaussian + DoG pyramid (k_sift_upsample + k_sift_blur<0..5> over all octaves), 256 N bytes per frame "
                              "(SURVEY 8d); the kernels back to back on one stream between two CUDA events (bm_sift_pyramid_ms)"},
                "roofline_matcher": None if mt is None else {
                    "bound": "tensor", "achieved": mt[1] / 1e12 / (mt[0] / 1e3), "peak": float(peaks.get("bf16_tflops", 2250.0)), "unit": "TFLOP/s",
                    "frac": mt[1] / 1e12 / (mt[0] / 1e3) / float(peaks.get("bf16_tflops", 2250.0)), "flops_per_pair": mt[1], "ms_per_pair": mt[0],
                    "kernel": "SIFT matcher (k_l2_knn2_tc: tcgen05 bf16 x bf16 -> f32 in TMEM, + merge, ratio test, stable sort), 700 x 700 x 128: one "
                              "frame pair is 0.125 GFLOP -- latency bound by construction, the tensor pipe idles (SURVEY 8d)"},
                "cpu_baseline": rm["cpu_baseline"], "frames_ok": rm["frames_ok"], "event_ms_per_step": rm["event_ms_per_step"],
                "regions": rm["regions"],
                "finalize": {"what": "crop_black_areas(80, 30) + scale_to_screen of the final canvas (main.py:1647-1659) via bm_finalize, "
                                     "result copied to the host", "device_ms": finalize_ms, "out_shape": list(final_img.shape),
                             "cpu_ms": finalize_cpu_ms,
                             "mosaic_jpg": {"what": "the same plus the JPEG file of main.py:1664-1665 encoded on the device (bm_finalize_jpeg): only the "
                                                    "compressed file is copied to the host; cpu_ms = cv2.imencode of the scaled image alone",
                                            "device_ms": finalize_jpeg_ms, "bytes": len(jpeg_bytes), "cpu_ms": jpeg_cpu_ms,
                                            "identical_to_cv2": jpeg_same}},
                "preview": {"what": "400 x 300 RGB progress thumbnail of the live canvas (gui.py:143-158: cvtColor + Pillow bicubic resize) via "
                                    "bm_preview, result copied to the host; cpu_ms excludes the full-canvas D2H the reference path would need",
                            "device_ms": preview_ms, "out_shape": list(thumb.shape), "cpu_ms": preview_cpu_ms},
                "host_cores_per_rank": len(my_cores) if my_cores else None}
        for det in dets[1:]:
            r = record(det)
            line[det] = {"metric": f"mosaic frames/sec at {w}x{h}, detector={det} (same sweep, same legs as the headline)", "unit": "frames/s", **r}
        if modes is not None:
            line["modes"] = modes
        if clip is not None:
            line["clip"] = clip
        if longrec is not None:
            line["long_run"] = longrec
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
