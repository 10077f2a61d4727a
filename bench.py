#!/usr/bin/env python
"""bench.py -- dgrad -> mesh frames/s on FLAME (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--sentences S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): configs[1] of BASELINE.json -- PCA-coefficient decode + reconstruction of
240-frame sentences (4 s at 60 fps) on the FLAME template with the default mask (5023 v / 9976 tris, 3762
constrained vertices, 2601 active triangles), random PCA basis (K = 85 scale + 180 rotation).  One step =
one batch of S sentences per GPU (default 315 -> 75 600 frames), i.e. the kernels of the path:
K1 decode -> K2 assembly -> K3 solve -> K5 output (transpose + base + constrained vertices).  Frames are sharded over ranks with no
data-path collective ("weak": per-GPU batch fixed).

  value     frames/s, inputs (coefficients, basis, factor) resident in HBM, CUDA-event timed, max over ranks
  e2e       same through the host-buffer C-ABI call (pinned host coefficients in, host vertices out)
  dgrad_resident  the same frames with the decoded dgrad [N, 9976*9] fp32 already in HBM (K2+K3 only):
            the literal "dgrad -> mesh" number SURVEY 8(d)'s 153 912 B/frame figure refers to
  roofline  dominant kernel, live CUDA-event time from the library's per-stage events; roofline_kernels lists all four
  cpu_baseline  the UNMODIFIED reference solver (oracle/_ref, all host threads) on a bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "sdfa-2019_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FRAMES_PER_SENTENCE = 240
N_VERTS, N_TRIS, N_FREE, N_ACTIVE = 5023, 9976, 1261, 2601
BYTES_PATH = 36 * N_ACTIVE + 12 * N_VERTS          # 153 912 B/frame, SURVEY 8(d)
BYTES_SOLVE = 2 * 12 * N_FREE                      # 30 264 B/frame: rhs in + solution out
BYTES_ASSEMBLY = 36 * N_ACTIVE + 12 * N_FREE       # dgrad of the active triangles in + rhs out
BYTES_OUTPUT = 12 * N_FREE + 12 * N_VERTS          # solved rows in + vertices out
DECODE_FLOP = 2 * (N_ACTIVE * 6 * 85 + N_ACTIVE * 3 * 180)   # 5 462 100 useful FLOP/frame


def ncu_traffic(kernel, n_frames):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/r1_traffic.json holds
    bytes per frame measured at 15 360 frames per launch, both decode launches added up), scaled to this launch;
    None if absent."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        with open(path) as fp:
            per_frame = json.load(fp)["dram_bytes_per_frame"][kernel]
        return per_frame * n_frames
    except Exception:
        return None


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


def workload(n_frames, seed):
    from deformation import workloads as W
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

    def step(self, xs, xr):
        torch = self.torch
        n = len(xs)
        s = torch.nn.functional.linear(torch.from_numpy(xs), self.cs, self.ms)
        r = torch.nn.functional.linear(torch.from_numpy(xr), self.cr, self.mr)
        dg = torch.cat((s.view(n, -1, 6), r.view(n, -1, 3)), dim=-1).view(n, -1).numpy()
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = FRAMES_PER_SENTENCE * 4
    V, F, nfv, pca, xs, xr = workload(n, seed=2)
    ref = CpuReference(V, F, nfv, pca, os.cpu_count() or 1)
    value, dt = ref.fps(xs, xr, steps=args.steps, warmup=args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": "dgrad->mesh frames/s (FLAME 5023v)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: PCA decode + reconstruction, FLAME default mask, bounded sample of "
                               f"{n} frames (4 sentences) per step on the host cores"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
                         "sample": f"{n} frames per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import deformation as D
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
    n = args.sentences * FRAMES_PER_SENTENCE
    V, F, nfv, pca, xs, xr = workload(n, seed=2 + rank)
    rec = D.Reconstructor(V, F, cnsts=nfv, device=local)
    rec.set_pca(*pca)
    xs_d, xr_d = torch.from_numpy(xs).to(dev), torch.from_numpy(xr).to(dev)
    out = torch.empty((n, N_VERTS, 3), dtype=torch.float32, device=dev)
    dgrad = rec.decode_dgrad(xs_d, xr_d)            # [n, 89784] fp32 resident dgrad for the K2+K3 leg
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    launches0 = D.lib.sdfa_launch_count()
    with ClockSampler(local) as clocks:
        ms_step = timed(lambda: rec.decode_and_get_mesh(xs_d, xr_d, out=out), args.steps, args.warmup)
    launches = (D.lib.sdfa_launch_count() - launches0) // max(1, args.steps + args.warmup) * args.steps
    ms_dgrad = timed(lambda: rec.get_mesh_batch(dgrad, out=out), args.steps, args.warmup)

    # end to end through the host-buffer call: pinned host coefficients in, host vertices out
    xs_h, xr_h = torch.from_numpy(xs).pin_memory(), torch.from_numpy(xr).pin_memory()
    out_h = torch.empty((n, N_VERTS, 3), dtype=torch.float32).pin_memory()
    xs_n, xr_n, out_n = xs_h.numpy(), xr_h.numpy(), out_h.numpy()

    def e2e_step():
        rec.decode_and_get_mesh(xs_n, xr_n, out=out_n)
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # per-kernel device time (library's own CUDA events on the launch stream), separate pass
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
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    total = n * world
    dom = max(("solve_ms", "assembly_ms", "decode_ms", "output_ms"), key=lambda k: stage[k])
    if dom == "decode_ms":
        # issued MMA work: 3 TF32 products per K step, K padded to 32-wide blocks (the means ride in one of the pad
        # columns), rows padded to 256-row tiles; peak = dense TF32 = half of the measured dense bf16 rate
        rows_s, rows_r = -(-6 * rec.n_active // 256) * 256, -(-3 * rec.n_active // 256) * 256
        issued = 3 * 2 * (rows_s * 96 + rows_r * 192)
        ach = issued * n / (stage[dom] * 1e-3) / 1e12
        peak = peaks["bf16_tflops"] / 2
        roof = {"kernel": "k_decode_tc (K1)", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": ncu_traffic("k_decode_tc", n),
                "issued_flop_per_frame": issued, "useful_flop_per_frame": DECODE_FLOP,
                "useful_tflops": DECODE_FLOP * n / (stage[dom] * 1e-3) / 1e12,
                "note": "kind::tf32 3xTF32; peak = measured dense bf16 / 2 (no TF32 entry in MEASURED_PEAKS.json); "
                        "the kernel also writes the 94 KB/frame compact dgrad"}
    else:
        b, kname = {"solve_ms": (BYTES_SOLVE, "k_solve_tc" if rec.debug("ts_stats")[0] else "k_solve"),
                    "assembly_ms": (BYTES_ASSEMBLY, "k_assemble"), "output_ms": (BYTES_OUTPUT, "k_output")}[dom]
        ach = b * n / (stage[dom] * 1e-3) / 1e9
        tag = {"solve_ms": " (K3)", "assembly_ms": " (K2)", "output_ms": " (K5)"}[dom]
        roof = {"kernel": kname + tag, "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic(kname, n),
                "algorithmic_bytes_per_frame": b, "algorithmic_bytes_per_launch": b * n}
    # every kernel of the path against its own algorithmic bytes (decode: the compact dgrad it must write)
    per_kernel = {}
    for key, nm, by in (("decode_ms", "k_decode_tc", 36 * N_ACTIVE), ("assembly_ms", "k_assemble", BYTES_ASSEMBLY),
                        ("solve_ms", "k_solve", BYTES_SOLVE), ("output_ms", "k_output", BYTES_OUTPUT)):
        gbs = by * n / (stage[key] * 1e-3) / 1e9
        per_kernel[nm] = {"ms": stage[key], "algorithmic_bytes_per_frame": by, "achieved_gbs": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"]}
    roof["peak_source"] = peak_src
    path_gbs = BYTES_PATH * n / (ms_dgrad * 1e-3) / 1e9
    result = {
        "metric": "dgrad->mesh frames/s (FLAME 5023v)", "value": total / (ms_step * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: PCA-coefficient decode + reconstruction, {args.sentences} sentences x 240 "
                               f"frames per GPU per step, FLAME 5023v/9976t default mask (1261 unknowns), random PCA "
                               "basis K=85+180", "frames_per_step_per_gpu": n, "parallelism": f"frames sharded x{world}",
                   "l2": "flushed between timed iterations (256 MiB memset) and inputs > L2"},
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(xs.nbytes + xr.nbytes),
                "d2h_bytes_per_step": int(n * N_VERTS * 12), "api": "deformation.Reconstructor.decode_and_get_mesh(numpy) "
                "-> sdfa_decode_reconstruct_host"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roof,
        "roofline_path": {"what": "K2+K3+fill with dgrad resident in HBM, 153 912 algorithmic B/frame (SURVEY 8d)",
                          "achieved": path_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": path_gbs / peaks["hbm_gbs"]},
        "dgrad_resident": {"value": total / (ms_dgrad * 1e-3), "unit": "frames/s", "ms_per_step": ms_dgrad},
        "kernel_ms_per_step": stage,
        "roofline_kernels": per_kernel,
    }
    if not args.no_cpu_baseline:
        ref = CpuReference(V, F, nfv, pca, os.cpu_count() or 1)
        sample = FRAMES_PER_SENTENCE * 8
        fps, _ = ref.fps(xs[:sample], xr[:sample], steps=3, warmup=1)
        result["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
                                  "sample": f"3 x {sample} frames (8 sentences), torch fp32 F.linear decode + reference "
                                            "get_mesh, one solver per host thread"}
    print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
