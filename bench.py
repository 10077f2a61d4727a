#!/usr/bin/env python
"""bench.py -- dgrad -> mesh frames/s on FLAME (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--sentences S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): configs[1] of BASELINE.json -- PCA-coefficient decode + reconstruction of
240-frame sentences (4 s at 60 fps) on the FLAME template with the default mask (5023 v / 9976 tris, 3762
constrained vertices, 2601 active triangles), random PCA basis (K = 85 scale + 180 rotation).  One step =
one batch of S sentences per GPU (default 315 -> 75 600 frames), i.e. the kernels of the path:
K1 decode -> K2 assembly -> K3 solve -> K5 output (transpose + base + constrained vertices).  Frames are sharded over
ranks with no data-path collective ("weak": per-GPU batch fixed).

  value     frames/s, inputs (coefficients, basis, factor) resident in HBM, CUDA-event timed, max over ranks
  e2e       same through the host-buffer C-ABI call (pinned host coefficients in, host vertices out)
  dgrad_resident  the same frames with the decoded dgrad [N, 9976*9] fp32 already in HBM (K2+K3+K5 only):
            the literal "dgrad -> mesh" number SURVEY 8(d)'s 153 912 B/frame figure refers to
  parity    max |dv| of sampled frames of the TIMED output against the compiled reference; non-zero exit on failure
  roofline  dominant kernel per SURVEY 8(d): algorithmic bytes (or useful FLOP) per launch / live CUDA-event time
  gather    (N > 1) the same step followed by the NCCL gather of the vertex buffers (north star): free rows all-gathered
            on a side stream chunk by chunk under the next chunk's kernels
  cpu_baseline  the UNMODIFIED reference solver (oracle/_ref, all host threads) on a bounded sample
"""
import argparse
import importlib.util
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sdfa-2019_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FRAMES_PER_SENTENCE = 240
N_VERTS, N_TRIS, N_FREE, N_ACTIVE = 5023, 9976, 1261, 2601
BYTES_PATH = 36 * N_ACTIVE + 12 * N_VERTS          # 153 912 B/frame, SURVEY 8(d)
BYTES_SOLVE = 2 * 12 * N_FREE                      # 30 264 B/frame: rhs in + solution out
BYTES_ASSEMBLY = 36 * N_ACTIVE + 12 * N_FREE       # dgrad of the active triangles in + rhs out
BYTES_OUTPUT = 12 * N_FREE + 12 * N_VERTS          # solved rows in + vertices out
BYTES_DECODE_OUT = 36 * N_ACTIVE                   # compact dgrad the unfused decode must write
DECODE_FLOP = 2 * (N_ACTIVE * 6 * 85 + N_ACTIVE * 3 * 180)   # 5 462 100 useful FLOP/frame, SURVEY 8(d)
METRIC = "dgrad->mesh frames/s (FLAME 5023v)"


def load_workloads():
    """deformation/workloads.py by path: pure numpy helpers, WITHOUT importing the ``deformation`` package (which loads
    libsdfa_b200.so) -- the reference arm must not map the product library."""
    spec = importlib.util.spec_from_file_location("sdfa_workloads", os.path.join(PKG, "deformation", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bench_config(sentences, world):
    n = sentences * FRAMES_PER_SENTENCE
    return {"workload": f"configs[1]: PCA-coefficient decode + reconstruction, {sentences} sentences x 240 "
                        f"frames per GPU per step, FLAME 5023v/9976t default mask (1261 unknowns), random PCA "
                        "basis K=85+180", "frames_per_step_per_gpu": n, "parallelism": f"frames sharded x{world}",
            "l2": "flushed between timed iterations (256 MiB memset) and inputs > L2"}


def traffic_table():
    """DRAM bytes per frame per kernel from the committed ncu --set full capture of this round (profiles/), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fp:
                return json.load(fp), name
        except Exception:
            continue
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            return json.load(fp), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi style clock/throttle sampling during the timed region (via NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index):
    """Run this rank (and first-touch its pinned staging buffers) on the host cores next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def workload(W, n_frames, seed):
    V, F, nfv, nft = W.load_flame()
    cs, ms, cr, mr = W.random_pca(len(F), seed=1, zero_tris=nft)
    xs, xr = W.random_coeffs(n_frames, seed=seed)
    return V, F, nfv, (cs, ms, cr, mr), xs, xr


class CpuReference:
    """Reference path on the host: torch fp32 F.linear x2 + cat (output_module.py:115-116, model.py:246-257)
    and the unmodified reference get_mesh per frame, one solver instance per host thread.  set_target is
    done once, outside the timed step, like the GPU arm's sdfa_create."""

    def __init__(self, V, F, nfv, pca, n_threads):
        import torch
        from oracle import ref_loader
        self.torch = torch
        torch.set_num_threads(max(1, n_threads))     # torchrun pins OMP_NUM_THREADS=1; the reference arm may use every core
        self.cs, self.ms, self.cr, self.mr = (torch.from_numpy(a) for a in pca)
        self.C = V[nfv]
        if ref_loader.ref_available():
            self.kind, self.cores = "reference", n_threads
            self.rs = ref_loader.RefSolver(n_threads)
            self.rs.set_target(V, F, cnsts=nfv)
        else:
            from oracle.dgrad_oracle import TriangleDeformationOracle
            self.kind, self.cores = "port", 1
            self.o = TriangleDeformationOracle()
            self.o.set_target(V, F, cnsts=nfv)

    def decode(self, xs, xr):
        torch = self.torch
        n = len(xs)
        s = torch.nn.functional.linear(torch.from_numpy(xs), self.cs, self.ms)
        r = torch.nn.functional.linear(torch.from_numpy(xr), self.cr, self.mr)
        return torch.cat((s.view(n, -1, 6), r.view(n, -1, 3)), dim=-1).view(n, -1).numpy()

    def step(self, xs, xr):
        dg = self.decode(xs, xr)
        if self.kind == "reference":
            out, _ = self.rs.get_mesh_batch(dg, self.C)
            return out
        return np.stack([self.o.get_mesh(d.astype(np.float64), vert_cnsts=self.C) for d in dg])

    def fps(self, xs, xr, steps=1, warmup=1):
        for _ in range(warmup):
            self.step(xs[:64], xr[:64])
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step(xs, xr)
        dt = time.perf_counter() - t0
        return len(xs) * steps / dt, dt


def run_reference(args):
    """The reference's own CPU implementation on the host cores; same metric / unit / config as the product arm, each
    step a bounded sample of that workload (4 sentences) so that the run ends within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = load_workloads()
    n = FRAMES_PER_SENTENCE * 4
    V, F, nfv, pca, xs, xr = workload(W, n, seed=2)
    ref = CpuReference(V, F, nfv, pca, os.cpu_count() or 1)
    value, dt = ref.fps(xs, xr, steps=args.steps, warmup=args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.sentences, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
                         "sample": f"bounded sample of the config's workload: {n} frames (4 sentences) per step x "
                                   f"{args.steps} steps, torch fp32 F.linear decode + reference get_mesh, one solver per host thread"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def parity_check(out_rows, sample_ids, xs, xr, ref, tol, free_ids=None):
    """Sampled frames of a timed output against the reference path (fp32 F.linear decode + reference get_mesh)."""
    want = ref.step(xs[sample_ids], xr[sample_ids])
    if free_ids is not None:
        want = want[:, free_ids]
    err = np.abs(out_rows - want).reshape(len(sample_ids), -1).max(axis=1)
    return float(err.max()), err


def _dist_setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return torch, dist, world, rank, local, dev, barrier, max_over_ranks


def run_config4(args):
    """configs[3]: audio -> mel (+ deltas) -> random-init temporal-attention network -> PCA coefficients -> K1..K5, the
    utterances sharded over the ranks; the network (plain PyTorch, not the product) is timed separately from the
    dgrad -> mesh path.  One step = every utterance of this rank once."""
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": "config 4: the reference network needs saber/librosa "
                              "(absent here); the reference arm exists for the headline config only"}))
        return
    torch, dist, world, rank, local, dev, barrier, max_over_ranks = _dist_setup()
    import deformation as D
    from deformation import frontend as FE, sharded, workloads as W
    V, F, nfv, nft = W.load_flame()
    pca = W.random_pca(len(F), seed=1, zero_tris=nft)
    rec = D.Reconstructor(V, F, cnsts=nfv, device=local)
    rec.set_pca(*pca)
    tol = 1e-6 * W.bbox_diag(V)
    lo, hi = sharded.shard_range(args.utterances, rank, world)
    n_utt, frames_per_utt = hi - lo, FRAMES_PER_SENTENCE
    n = n_utt * frames_per_utt
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    feats_mod, net = FE.MelFeatures().to(dev), FE.build_network(seed=0, device=dev, coeff_gain=8.0)
    audio = FE.band_limited_noise(n_utt, seconds=4.0, seed=100 + rank, device=dev)     # [n_utt, 32000], resident in HBM
    speaker = (torch.arange(n_utt, device=dev) + lo) % net.num_speakers
    xs = torch.empty((n, 85), dtype=torch.float32, device=dev)
    xr = torch.empty((n, 180), dtype=torch.float32, device=dev)
    out = torch.empty((n, N_VERTS, 3), dtype=torch.float32, device=dev)
    UB = 8                                                      # utterances per network batch (1920 windows, 0.75 GB of features)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step(n_batches=None):
        t_feat = t_net = 0.0
        marks = []
        with torch.no_grad():
            for bi, u0 in enumerate(range(0, n_utt, UB)):
                if n_batches is not None and bi >= n_batches:
                    break
                u1 = min(n_utt, u0 + UB)
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                f = feats_mod(audio[u0:u1], frames_per_utt)                           # [u, 240, 64, 128, 3]
                e1.record()
                spk = speaker[u0:u1].repeat_interleave(frames_per_utt)
                cs_, cr_ = net(f.reshape(-1, 64, 128, 3), spk)
                xs[u0 * frames_per_utt: u1 * frames_per_utt] = cs_
                xr[u0 * frames_per_utt: u1 * frames_per_utt] = cr_
                e2.record()
                marks.append((e0, e1, e2))
            m0, m1 = ev(), ev()
            m0.record()
            if n_batches is None:
                rec.decode_and_get_mesh(xs, xr, out=out)
            else:
                k = min(n, n_batches * UB * frames_per_utt)
                rec.decode_and_get_mesh(xs[:k], xr[:k], out=out[:k])
            m1.record()
        torch.cuda.synchronize()
        for e0, e1, e2 in marks:
            t_feat += e0.elapsed_time(e1)
            t_net += e1.elapsed_time(e2)
        return t_feat, t_net, m0.elapsed_time(m1)

    xs.zero_()
    xr.zero_()
    for _ in range(max(1, args.warmup)):
        step(n_batches=2)                                       # warm-up: cuDNN plans, allocator ...
        rec.decode_and_get_mesh(xs, xr, out=out)                # ... and the path's workspaces at the full batch size
    barrier()
    steps = max(1, min(args.steps, 3))
    launches0 = D.lib.sdfa_launch_count()
    tf = tn = tm = 0.0
    with ClockSampler(local) as clocks:
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            f_, n_, m_ = step()
            tf, tn, tm = tf + f_ / steps, tn + n_ / steps, tm + m_ / steps
        b.record()
        barrier()
    ms_step = max_over_ranks(a.elapsed_time(b) / steps)
    tf, tn, tm = max_over_ranks(tf), max_over_ranks(tn), max_over_ranks(tm)
    launches = (D.lib.sdfa_launch_count() - launches0) // steps
    ids = np.unique(np.linspace(0, n - 1, 12).astype(np.int64))
    sel = torch.from_numpy(ids).to(dev)
    got, cxs, cxr = out[sel].cpu().numpy(), xs[sel].cpu().numpy(), xr[sel].cpu().numpy()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ref = CpuReference(V, F, nfv, pca, os.cpu_count() or 1)
    worst, _ = parity_check(got, np.arange(len(ids)), cxs, cxr, ref, tol)
    total = args.utterances * frames_per_utt
    print(json.dumps({
        "metric": "audio->mesh frames/s (config 4, FLAME 5023v)", "value": total / (ms_step * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 (network: TF32 matmuls allowed)", "data": "synthetic",
        "config": {"workload": f"configs[3]: {args.utterances} utterances x 4 s band-limited noise at 8 kHz -> 240 windows "
                               "[64,128,3] (mel + delta + delta-delta) each -> random-init network of config/model/dgrad.py "
                               "-> coefficients [85]+[180] -> decode + reconstruction, FLAME default mask",
                   "utterances_per_gpu": n_utt, "frames_per_gpu": n, "parallelism": f"utterances sharded x{world}",
                   "network_batch_windows": UB * frames_per_utt},
        "stage_ms_per_step": {"features_ms": tf, "network_ms": tn, "dgrad_to_mesh_ms": tm},
        "dgrad_to_mesh": {"value": n * world / (tm * 1e-3), "unit": "frames/s",
                          "what": "decode_and_get_mesh on the network's coefficients, device resident, same step"},
        "network": {"value": n * world / ((tf + tn) * 1e-3), "unit": "frames/s", "what": "features + network, plain PyTorch (not the product)"},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "parity": {"max_abs_m": worst, "tol": tol, "frames": [int(i) for i in ids], "checker": ref.kind,
                   "what": "sampled frames of the timed output vs fp32 F.linear decode + reference get_mesh on the network's own coefficients",
                   "ok": bool(worst <= tol)},
    }))
    if dist is not None:
        dist.destroy_process_group()
    if worst > tol:
        sys.exit(1)


def run_config5(args):
    """configs[4]: twice-subdivided FLAME (79 936 v / 159 616 tris, 20 653 unknowns, propagated mask), synthetic iid dgrad
    resident in HBM in the reference layout, frames sharded over the ranks (large-factor solve: the SIMT sweeps)."""
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        W = load_workloads()
        from oracle import ref_loader
        V, F, c = W.flame_sub2()
        nt = min(4, os.cpu_count() or 1)
        rs = ref_loader.RefSolver(nt)
        rs.set_target(V, F, cnsts=c)
        dg = W.iid_dgrad(2 * nt, len(F), sigma=0.01, seed=5)
        rs.get_mesh_batch(dg[:nt], V[c])
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps)):
            rs.get_mesh_batch(dg, V[c])
        dt = (time.perf_counter() - t0) / max(1, args.steps)
        v = len(dg) / dt
        print(json.dumps({"impl": "reference", "metric": "dgrad->mesh frames/s (subdivided FLAME 79936v)", "value": v,
                          "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": "configs[4] (bounded sample)"},
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": nt, "kind": "reference",
                                           "sample": f"{len(dg)} frames per step on {nt} host threads, one reference solver each"},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    torch, dist, world, rank, local, dev, barrier, max_over_ranks = _dist_setup()
    import deformation as D
    from deformation import sharded, workloads as W
    V, F, c = W.flame_sub2()
    rec = D.Reconstructor(V, F, cnsts=c, device=local)
    tol = 1e-6 * W.bbox_diag(V)
    lo, hi = sharded.shard_range(args.frames, rank, world)
    n = hi - lo
    width = len(F) * 9
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    dgrad = torch.empty((n, width), dtype=torch.float32, device=dev)         # 5.75 MB per frame: 57 GB for 10 000 frames
    for a in range(0, n, 256):
        dgrad[a:a + 256].normal_(0.0, 0.01, generator=g)
    out = torch.empty((n, len(V), 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    steps = max(1, args.steps)
    for _ in range(max(1, args.warmup)):
        rec.get_mesh_batch(dgrad, out=out)
    barrier()
    launches0 = D.lib.sdfa_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with ClockSampler(local) as clocks:
        for a, b in evs:
            flush.zero_()
            a.record()
            rec.get_mesh_batch(dgrad, out=out)
            b.record()
        barrier()
    ms_step = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps)
    launches = (D.lib.sdfa_launch_count() - launches0) // steps
    rec.set_timing(True)
    stage = {"assembly_ms": 0.0, "solve_ms": 0.0, "output_ms": 0.0}
    for i in range(3):
        rec.get_mesh_batch(dgrad, out=out)
        if i:
            t = rec.last_timing()
            for k in stage:
                stage[k] += t[k] / 2
    rec.set_timing(False)
    stage = {k: max_over_ranks(v) for k, v in stage.items()}
    ids = np.unique(np.array([0, 1, n // 2, n - 1]))
    sel = torch.from_numpy(ids).to(dev)
    got, dg_s = out[sel].cpu().numpy(), dgrad[sel].cpu().numpy()
    # end to end through the host-buffer call on a bounded slice (pinned host dgrad in, host vertices out)
    ne = min(n, 512)
    dg_h = dgrad[:ne].cpu().pin_memory().numpy()
    out_h = torch.empty((ne, len(V), 3), dtype=torch.float32).pin_memory().numpy()
    rec.get_mesh_batch(dg_h, out=out_h)
    barrier()
    t0 = time.perf_counter()
    rec.get_mesh_batch(dg_h, out=out_h)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    from oracle import ref_loader
    from oracle.dgrad_oracle import TriangleDeformationOracle
    chk = ref_loader.RefSolver(1) if ref_loader.ref_available() else TriangleDeformationOracle()
    chk.set_target(V, F, cnsts=c)
    worst = 0.0
    for i in range(len(ids)):
        want = chk.get_mesh(dg_s[i].astype(np.float64), vert_cnsts=V[c])
        worst = max(worst, float(np.abs(got[i] - want).max()))
    peaks, peak_src = measured_peaks()
    n_active = rec.n_active
    bytes_path = 36 * n_active + 12 * len(V)                                  # 2 457 408 B/frame, SURVEY 8(d)
    bytes_solve = 2 * 12 * rec.n_free
    total = args.frames
    dom = max(stage, key=lambda k: stage[k])
    kb = {"assembly_ms": ("k_assemble_gather (K2)", 36 * n_active + 12 * rec.n_free), "solve_ms": ("k_solve (K3, SIMT sweeps)", bytes_solve),
          "output_ms": ("k_output (K5)", 12 * rec.n_free + 12 * len(V))}
    path_gbs = bytes_path * total / (ms_step * 1e-3) / 1e9
    ach = kb[dom][1] * n / (stage[dom] * 1e-3) / 1e9
    result = {
        "metric": "dgrad->mesh frames/s (subdivided FLAME 79936v)", "value": total / (ms_step * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[4]: FLAME subdivided twice ({len(V)} v / {len(F)} tris, {rec.n_free} unknowns, "
                               f"{n_active} active triangles, nnz(L) {rec.nnz_l}), {args.frames} frames iid sigma 0.01 dgrad resident "
                               "in HBM (device RNG), reference layout", "frames_per_gpu": n,
                   "parallelism": f"frames sharded x{world}", "l2": "flushed between timed iterations; input 5.7 MB per frame"},
        "e2e": {"value": ne * world / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(ne * width * 4),
                "d2h_bytes_per_step": int(ne * len(V) * 12), "ms_per_step": e2e_ms,
                "api": f"get_mesh_batch(numpy) -> sdfa_reconstruct_host on a bounded slice of {ne} frames per GPU, pinned buffers"},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "parity": {"max_abs_m": worst, "tol": tol, "frames": [int(i) for i in ids],
                   "checker": "reference" if ref_loader.ref_available() else "port", "ok": bool(worst <= tol)},
        "roofline": {"kernel": kb[dom][0], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes_per_frame": kb[dom][1], "peak_source": peak_src},
        "roofline_path": {"what": f"whole call, {bytes_path} algorithmic B/frame (SURVEY 8d)", "achieved": path_gbs,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": path_gbs / peaks["hbm_gbs"]},
        "kernel_ms_per_step": stage, "frames_per_solve_tile": int(rec.debug("stats")[14]),
    }
    print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()
    if worst > tol:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sentences", type=int, default=315,
                    help="240-frame sentences per step per GPU (315 -> 75 600 frames = 1773 solve tiles of 128 columns "
                         "= 11.98 per SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the NCCL gather leg at N > 1")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="BASELINE.json configs[] entry (1-based): 2 = the headline workload (default), 4 = audio -> mel -> "
                         "network -> mesh end to end, 5 = subdivided FLAME (large-factor solve)")
    ap.add_argument("--utterances", type=int, default=1000, help="config 4: 4 s utterances in total, sharded over the ranks")
    ap.add_argument("--frames", type=int, default=10000, help="config 5: frames in total, sharded over the ranks")
    args = ap.parse_args()
    if args.config == 4:
        return run_config4(args)
    if args.config == 5:
        return run_config5(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    cpus = bind_to_gpu_numa_node(local)
    import deformation as D
    from deformation import sharded, workloads as W
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.sentences * FRAMES_PER_SENTENCE
    V, F, nfv, pca, xs, xr = workload(W, n, seed=2 + rank)
    rec = D.Reconstructor(V, F, cnsts=nfv, device=local)
    rec.set_pca(*pca)
    tol = 1e-6 * W.bbox_diag(V)
    xs_d, xr_d = torch.from_numpy(xs).to(dev), torch.from_numpy(xr).to(dev)
    out = torch.empty((n, N_VERTS, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()                           # L2 flush between timed iterations (outside the events)
            a.record()
            fn()
            b.record()
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / steps

    t_start = time.time()

    def trace(leg):                                 # progress on stderr (every rank): which leg a failed run was in
        print(f"bench.py[rank {rank}] +{time.time() - t_start:6.1f}s {leg}", file=sys.stderr, flush=True)

    # ---- the timed step, then parity of what it just wrote
    trace("headline: decode + reconstruction")
    launches0 = D.lib.sdfa_launch_count()
    with ClockSampler(local) as clocks:
        ms_step = timed(lambda: rec.decode_and_get_mesh(xs_d, xr_d, out=out), args.steps, args.warmup)
    launches = (D.lib.sdfa_launch_count() - launches0) // max(1, args.steps + args.warmup) * args.steps
    # frames: first / last, 128-frame tile boundaries, and a spread over the batch (= over the persistent CTAs' tile ranges)
    sample_ids = np.unique(np.clip(np.concatenate([[0, 1, 127, 128, 129, n - 129, n - 128, n - 1],
                                                   np.linspace(0, n - 1, 24).astype(np.int64)]), 0, n - 1))
    out_samples = out[torch.from_numpy(sample_ids).to(dev)].cpu().numpy()

    # ---- the literal metric: dgrad resident in HBM (gather assembly + solve + output)
    trace("dgrad resident")
    dgrad = rec.decode_dgrad(xs_d, xr_d)            # [n, 89784] fp32
    ms_dgrad = timed(lambda: rec.get_mesh_batch(dgrad, out=out), args.steps, args.warmup)
    out_samples_dgrad = out[torch.from_numpy(sample_ids).to(dev)].cpu().numpy()
    del dgrad
    torch.cuda.empty_cache()

    # ---- configs[1] literally: ONE 240-frame sentence per call (latency of the call, device resident, CUDA events)
    one = FRAMES_PER_SENTENCE
    out_one = torch.empty((one, N_VERTS, 3), dtype=torch.float32, device=dev)
    ms_sentence = timed(lambda: rec.decode_and_get_mesh(xs_d[:one], xr_d[:one], out=out_one), max(args.steps, 20), args.warmup)

    # ---- per-kernel device time (library's own CUDA events on the launch stream), separate pass
    trace("per-kernel times")
    rec.set_timing(True)
    stage = {"decode_ms": 0.0, "assembly_ms": 0.0, "solve_ms": 0.0, "output_ms": 0.0}
    reps = max(3, min(args.steps, 10))
    for i in range(reps + 1):
        flush.zero_()
        rec.decode_and_get_mesh(xs_d, xr_d, out=out)
        if i:                                       # first one is a warm-up
            for k, v in rec.last_timing().items():
                stage[k] += v / reps
    rec.set_timing(False)

    # ---- end to end through the host-buffer calls: host coefficients in, host vertices out, wall clock
    trace("end to end (host buffers)")
    def e2e(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)

    xs_n, xr_n = torch.from_numpy(xs).pin_memory().numpy(), torch.from_numpy(xr).pin_memory().numpy()
    free_h = torch.empty((n, N_FREE, 3), dtype=torch.float32).pin_memory()
    free_n = free_h.numpy()
    e2e_free_ms = e2e(lambda: rec.decode_and_get_mesh(xs_n, xr_n, out=free_n, free_only=True))
    free_samples = free_n[sample_ids].copy()
    # what the host link gives for the same bytes with nothing else going on: plain pinned device->host copies of the
    # free rows, all ranks at once (the e2e number cannot exceed it)
    free_d = torch.empty((n, N_FREE, 3), dtype=torch.float32, device=dev)
    d2h_ms = e2e(lambda: (free_h.copy_(free_d, non_blocking=True), torch.cuda.synchronize()))
    del free_h, free_n, free_d
    full_h = torch.empty((n, N_VERTS, 3), dtype=torch.float32).pin_memory()
    full_n = full_h.numpy()
    e2e_full_ms = e2e(lambda: rec.decode_and_get_mesh(xs_n, xr_n, out=full_n))
    del full_h, full_n
    # what a caller with ordinary (pageable) numpy arrays gets -- fewer steps, it is slow
    page_out = np.empty((n, N_VERTS, 3), dtype=np.float32)
    page_steps = max(1, min(3, args.steps))
    rec.decode_and_get_mesh(xs, xr, out=page_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(page_steps):
        rec.decode_and_get_mesh(xs, xr, out=page_out)
    barrier()
    e2e_page_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / page_steps)
    del page_out

    # ---- N > 1: the step followed by the gather of the vertex buffers (free rows over NCCL, side stream, chunked)
    gather = None
    if world > 1 and not args.no_gather:
        gather = {}
        del out
        torch.cuda.empty_cache()
        row_b = N_FREE * 12
        for key, mode, expand in (("all_gather_free_rows", "all", False), ("all_gather_expanded", "all", True),
                                  ("gather_to_rank0_free_rows", "root", False)):
            trace("gather leg " + key)
            pipe = sharded.GatherPipeline(rec, chunk_frames=148 * 128, mode=mode, dst=0, expand=expand)
            rows = N_VERTS if expand else N_FREE
            g_out = (torch.empty((world * n, rows, 3), dtype=torch.float32, device=dev)
                     if (mode == "all" or rank == 0) else None)

            def step(pipe=pipe, g_out=g_out):
                pipe.run(lambda a, b, out: rec.decode_and_get_mesh(a, b, out=out, free_only=True), [xs_d, xr_d], out=g_out)
            ms = timed(step, max(3, args.steps // 2), 2)
            recv = (world - 1) * n * row_b            # bytes arriving at a receiving GPU per step
            gather[key] = {"value": world * n / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                           "nvlink_in_gbs_per_receiver": recv / (ms * 1e-3) / 1e9, "bytes_per_frame_on_wire": row_b}
            if key == "all_gather_expanded":
                # the gathered + expanded frames of rank r are this rank's timed output when r == rank
                mine = g_out[rank * n + torch.from_numpy(sample_ids).to(dev)].cpu().numpy()
                gather[key]["equals_local_output"] = bool(np.array_equal(mine, out_samples))
            del g_out, pipe, step                  # (the closure's defaults hold both buffers as well)
            torch.cuda.empty_cache()
        # the same gather with our own data path: symmetric memory + peer-to-peer pushes on a side stream (no NCCL kernel)
        for key, mode in (("p2p_all_gather_free_rows", "all"), ("p2p_gather_to_rank0_free_rows", "root")):
            trace("gather leg " + key)
            try:
                pg = sharded.PeerGather(rec, n, chunk_frames=148 * 128, mode=mode, dst=0)
            except Exception as ex:                      # no peer mapping on this box: the NCCL numbers above stand
                gather[key] = {"unavailable": str(ex)[:200]}
                barrier()                                # (the same barrier the other branch ends with)
                continue

            def step(pg=pg):
                pg.run(lambda a, b, out: rec.decode_and_get_mesh(a, b, out=out, free_only=True), [xs_d, xr_d])
            ms = timed(step, max(3, args.steps // 2), 2)
            recv = (world - 1) * n * row_b if (mode == "all" or rank == 0) else 0
            gather[key] = {"value": world * n / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                           "nvlink_in_gbs_per_receiver": (world - 1) * n * row_b / (ms * 1e-3) / 1e9,
                           "nvlink_out_gbs_per_sender": pg.bytes_pushed_per_run / (ms * 1e-3) / 1e9, "bytes_per_frame_on_wire": row_b}
            if mode == "all":
                mine = pg.buf[rank * n + torch.from_numpy(sample_ids).to(dev)].cpu().numpy()
                gather[key]["equals_local_output"] = bool(np.array_equal(mine, out_samples[:, rec.free_vertices]))
            del pg, step                           # the closure's default holds the symmetric buffer and its peer mappings
            gc.collect()
            torch.cuda.empty_cache()
            barrier()                              # every rank has let go of the symmetric memory before anyone moves on (or exits)
        gather["limiter"] = ("NVLink: (N-1)/N of every frame's free rows (15 132 B) leave each sender N-1 times and arrive at each "
                             "receiver; at N = 8 the all-gather needs 7 x 1.14 GB per GPU and step in each direction "
                             "(900 GB/s per direction => 8.9 ms against 5.3 ms of kernels); the kernels overlap on the main stream")

    trace("multi-rank part done")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rank 0: parity of the timed outputs, rooflines, CPU baseline
    ref = CpuReference(V, F, nfv, pca, os.cpu_count() or 1)
    free_ids = rec.free_vertices
    worst, _ = parity_check(out_samples, sample_ids, xs, xr, ref, tol)
    worst_dgrad, _ = parity_check(out_samples_dgrad, sample_ids, xs, xr, ref, tol)
    worst_free, _ = parity_check(free_samples, sample_ids, xs, xr, ref, tol, free_ids)
    cnst_ok = bool(np.array_equal(out_samples[:, nfv], np.broadcast_to(V[nfv], (len(sample_ids),) + V[nfv].shape)))
    parity = {"max_abs_m": max(worst, worst_dgrad, worst_free), "tol": tol, "frames": [int(i) for i in sample_ids],
              "checker": ref.kind, "max_abs_m_decode_path": worst, "max_abs_m_dgrad_path": worst_dgrad,
              "max_abs_m_e2e_free_rows": worst_free, "constrained_rows_bit_equal": cnst_ok,
              "ok": bool(max(worst, worst_dgrad, worst_free) <= tol and cnst_ok)}

    peaks, peak_src = measured_peaks()
    traffic, traffic_src = traffic_table()

    def ncu_traffic(kernel):
        try:
            return traffic["dram_bytes_per_frame"][kernel] * n
        except Exception:
            return None

    # dense TF32 matmul measured in this run (no TF32 entry in MEASURED_PEAKS.json)
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev)
    b = torch.randn(8192, 8192, device=dev)
    for _ in range(3):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    tf32_tflops = 10 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    a, b = a.half(), b.half()
    for _ in range(3):
        a @ b
    e0.record()
    for _ in range(10):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    f16_tflops = 10 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b
    decode_kernel = "k_decode_tc16" if rec.debug("decode_kind")[0] == 16 else "k_decode_tc"
    tensor_peak = f16_tflops if decode_kernel == "k_decode_tc16" else tf32_tflops

    total = n * world
    dom = max(("solve_ms", "assembly_ms", "decode_ms", "output_ms"), key=lambda k: stage[k])
    solver_kernel = "k_solve_tc" if rec.debug("ts_stats")[0] else "k_solve"
    if dom == "decode_ms":
        rows_s, rows_r = -(-6 * rec.n_active // 256) * 256, -(-3 * rec.n_active // 256) * 256
        issued = 3 * 2 * (rows_s * 96 + rows_r * 192)     # 3 split products per K step, K and rows padded to tiles
        ach = DECODE_FLOP * n / (stage[dom] * 1e-3) / 1e12
        roof = {"kernel": decode_kernel + " (K1)", "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": ach / tensor_peak, "traffic": ncu_traffic(decode_kernel),
                "useful_flop_per_frame": DECODE_FLOP, "issued_flop_per_frame": issued,
                "issued_frac": issued * n / (stage[dom] * 1e-3) / 1e12 / tensor_peak,
                "peak_source": "torch 8192^3 matmul in the kernel's operand type (fp16, or fp32 with allow_tf32), measured in this run",
                "note": "achieved = SURVEY 8(d) useful FLOP (active triangles, K = 85 / 180, one product); the kernel issues "
                        "three split products on padded tiles (issued_frac) and writes the 94 KB/frame compact dgrad"}
    else:
        b_, kname = {"solve_ms": (BYTES_SOLVE, solver_kernel), "assembly_ms": (BYTES_ASSEMBLY, "k_assemble"),
                     "output_ms": (BYTES_OUTPUT, "k_output2")}[dom]
        ach = b_ * n / (stage[dom] * 1e-3) / 1e9
        tag = {"solve_ms": " (K3)", "assembly_ms": " (K2)", "output_ms": " (K5)"}[dom]
        roof = {"kernel": kname + tag, "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic(kname),
                "algorithmic_bytes_per_frame": b_, "algorithmic_bytes_per_launch": b_ * n, "peak_source": peak_src}
    roof["traffic_source"] = traffic_src
    # every kernel of the path against its own algorithmic bytes (decode: useful FLOP against the TF32 peak as well)
    per_kernel = {}
    for key, nm, by in (("decode_ms", decode_kernel, BYTES_DECODE_OUT), ("assembly_ms", "k_assemble", BYTES_ASSEMBLY),
                        ("solve_ms", solver_kernel, BYTES_SOLVE), ("output_ms", "k_output2", BYTES_OUTPUT)):
        gbs = by * n / (stage[key] * 1e-3) / 1e9
        per_kernel[nm] = {"ms": stage[key], "algorithmic_bytes_per_frame": by, "achieved_gbs": gbs,
                          "frac_of_hbm": gbs / peaks["hbm_gbs"], "ncu_dram_bytes_per_launch": ncu_traffic(nm)}
    per_kernel[decode_kernel]["useful_tflops"] = DECODE_FLOP * n / (stage["decode_ms"] * 1e-3) / 1e12
    per_kernel[decode_kernel]["frac_of_tensor_peak"] = per_kernel[decode_kernel]["useful_tflops"] / tensor_peak
    per_kernel[decode_kernel]["tensor_peak_tflops"] = tensor_peak
    path_gbs = BYTES_PATH * n / (ms_dgrad * 1e-3) / 1e9
    result = {
        "metric": METRIC, "value": total / (ms_step * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.sentences, world),
        "e2e": {"value": total / (e2e_free_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(xs.nbytes + xr.nbytes),
                "d2h_bytes_per_step": int(n * N_FREE * 12), "ms_per_step": e2e_free_ms,
                "api": "deformation.Reconstructor.decode_and_get_mesh(numpy, free_only=True) -> "
                       "sdfa_decode_reconstruct_free_host: pinned host coefficients in, the 1261 free vertices of every frame "
                       "out (the 3762 constrained rows are the caller's own vert_cnsts constants)",
                "host_link_ceiling": {"value": total / (d2h_ms * 1e-3), "unit": "frames/s", "ms_per_step": d2h_ms,
                                      "aggregate_d2h_gbs": world * n * N_FREE * 12 / (d2h_ms * 1e-3) / 1e9,
                                      "what": "plain pinned cudaMemcpy D2H of the same free-row bytes, all ranks at once, no kernels"}},
        "e2e_full_layout": {"value": total / (e2e_full_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_full_ms,
                            "d2h_bytes_per_step": int(n * N_VERTS * 12),
                            "api": "decode_and_get_mesh(numpy) -> sdfa_decode_reconstruct_host, pinned buffers, all 5023 rows"},
        "e2e_pageable": {"value": total / (e2e_page_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_page_ms,
                         "api": "decode_and_get_mesh(numpy) with ordinary pageable arrays, all 5023 rows"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "parity": parity,
        "roofline": roof,
        "roofline_path": {"what": "K2+K3+K5 with dgrad resident in HBM, 153 912 algorithmic B/frame (SURVEY 8d)",
                          "achieved": path_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": path_gbs / peaks["hbm_gbs"]},
        "dgrad_resident": {"value": total / (ms_dgrad * 1e-3), "unit": "frames/s", "ms_per_step": ms_dgrad},
        "single_sentence": {"frames": one, "ms_per_call": ms_sentence, "value": one / (ms_sentence * 1e-3), "unit": "frames/s",
                            "what": "configs[1] as one call: decode + reconstruction of one 240-frame sentence, device resident, "
                                    "L2 flushed before every call"},
        "kernel_ms_per_step": stage,
        "roofline_kernels": per_kernel,
        "tf32_matmul_tflops_measured": tf32_tflops, "f16_matmul_tflops_measured": f16_tflops,
        "host_cpus_bound": len(cpus) if cpus else None,
    }
    if gather is not None:
        result["gather"] = gather
    if not args.no_cpu_baseline:
        sample = FRAMES_PER_SENTENCE * 8
        fps, _ = ref.fps(xs[:sample], xr[:sample], steps=3, warmup=1)
        result["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
                                  "sample": f"3 x {sample} frames (8 sentences), torch fp32 F.linear decode + reference "
                                            "get_mesh, one solver per host thread"}
    print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()
    if not parity["ok"]:
        print(f"bench.py: PARITY FAILED: {parity}", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
