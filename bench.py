#!/usr/bin/env python
"""Benchmark of the compression forward path (encode + rate), BASELINE.json metric:
images/s masked-ViT encode+rate.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload B64|B144|L256]

One "step" = one pass of the path (mask-select -> gather -> ViT encoder -> g_a/h_a/h_s/slice nets -> quantise ->
likelihoods -> bpp) over one batch of synthetic images.  Default workload = BASELINE.json configs[1]:
synthetic 224x224, batch 64 per GPU, ViT-B/16, K=64 kept patches (nearest valid value to mask ratio 0.75; the
reference needs sqrt(K) % 4 == 0), bf16 tensor-core operands, fp32 accumulate.
Images are sharded across ranks (weak scaling, weights replicated); the only collective is the 16-byte rate
all-reduce per step.

--impl reference: the reference's CPU implementation of the same path (the in-repo fp32 oracle; the reference module
itself cannot be imported - timm/compressai absent) on the host cores, on a bounded sample of the same workload.
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

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

WORKLOADS = {
    # name: (model kwargs, per-GPU batch, description)
    "B64": (dict(img_size=224, num_keep_patches=64), 64,
            "synthetic 224x224 batch 64/GPU, ViT-B/16, K=64 (67% masked; nearest valid to 0.75), bf16"),
    "B144": (dict(img_size=224, num_keep_patches=144), 64,
             "synthetic 224x224 batch 64/GPU, ViT-B/16, K=144 (shipped test.sh value), bf16"),
    "L256": (dict(img_size=512, encoder_embed_dim=1024, encoder_depth=24, encoder_num_heads=16, num_keep_patches=256), 32,
             "synthetic 512x512 batch 32/GPU, ViT-L/16, K=256 (75% masked), bf16"),
}
ALGO_GFLOP_PER_IMG = {"B64": 20.761, "B144": 46.636, "L256": 201.005}     # SURVEY 6.2 (reference-algorithmic)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(kwargs, batch, seed, n_rot):
    g = torch.Generator().manual_seed(seed)
    S = kwargs["img_size"]
    L = (S // 16) ** 2
    imgs = [torch.rand(batch, 3, S, S, generator=g) for _ in range(n_rot)]
    scores = [torch.rand(batch, L, generator=g) for _ in range(n_rot)]
    return imgs, scores


# ----------------------------------------------------------------------------------------------------------
def run_reference(args, kwargs, batch, desc):
    """The reference's CPU path (fp32 oracle == reference math) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_model
    from textmae_image_compression_b200 import PathConfig, make_state_dict
    cfg = PathConfig(**kwargs)
    sd = make_state_dict(cfg, seed=0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(batch, args.ref_sample)
    imgs, scores = make_inputs(kwargs, sample, 0, 1)
    for _ in range(max(args.warmup, 1) if args.warmup < 2 else 1):
        ref_model.forward_rate(sd, cfg, imgs[0], scores[0])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_model.forward_rate(sd, cfg, imgs[0], scores[0])
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "images/s masked-ViT encode+rate", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample": f"{sample} of {batch} images per step"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} images/step x {args.steps} steps, fp32 oracle (oracle/ref_model.py), "
                                   f"host python mask routine included"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_b200(args, kwargs, batch, desc, wl):
    import torch.distributed as dist
    from textmae_image_compression_b200 import MCM, PathConfig, make_state_dict
    from textmae_image_compression_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = PathConfig(**kwargs)
    sd = make_state_dict(cfg, seed=0)                         # replicated weights
    # `--streams S` pipelines S batches through S independent handles (own workspace, own CUDA stream): the serial
    # slice chain of one batch leaves most SMs idle, a second batch in flight fills them.  A step is still one batch.
    S = max(1, args.streams)
    models = []
    for _ in range(S):
        m_ = MCM(**kwargs, skip_dead_lrp=False, share_sm=S > 1)
        m_.load_state_dict(sd)
        m_.cuda().eval()
        models.append(m_)
    model = models[0]
    del sd
    n_rot = 8                                                 # rotating input batches: 8 x 38.5 MB > 126 MB L2
    imgs_h, scores_h = make_inputs(kwargs, batch, 1000 + rank, n_rot)
    imgs_d = [t.cuda() for t in imgs_h]
    scores_d = [t.cuda() for t in scores_h]
    for m_ in models:
        m_.reserve(batch)
    stream = torch.cuda.current_stream()
    side = [torch.cuda.Stream(device=dev) for _ in range(S)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(i):
        with torch.cuda.stream(side[i % S]):
            out = models[i % S](imgs_d[i % n_rot], scores_d[i % n_rot])
            if world > 1:
                D.aggregate_rate(out["rate_sums"])            # the path's only collective (16 bytes)
        return out

    def fork(ev):
        ev.record(stream)
        for s_ in side:
            s_.wait_event(ev)

    def join(ev):
        for s_ in side:
            stream.wait_stream(s_)
        ev.record(stream)

    # ---- device-resident throughput -------------------------------------------------------------------
    # every handle needs 3 forwards before it is in steady state (plain launches, graph capture, first replay) and its
    # stream's allocator pool is populated: warm up max(W, 3 S) steps so none of that lands in the timed region
    warm_eff = max(args.warmup, 3 * S)
    # the clock sampler (nvidia-smi -lms 100) starts BEFORE the warm-up: its process start / NVML initialisation stalls
    # the GPU for tens of ms, which must not fall into a 40 ms timed region; it then samples through both timed regions
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for i in range(S):                                        # first call per handle: lazy initialisation (plans, NCCL)
        step_dev(i)
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    for i in range(S, warm_eff):
        step_dev(i)
    torch.cuda.synchronize()
    # at least ~1 s under load before timing (clocks / power state settle).  The number of extra steps is agreed across
    # ranks (every step carries the 16-byte rate all-reduce when world > 1, so all ranks must run the same count).
    el = time.perf_counter() - t_warm
    extra = 0 if el >= 1.0 else int((1.0 - el) / max(el / max(warm_eff - S, 1), 1e-5) / S + 1) * S
    if world > 1:
        t_extra = torch.tensor([extra], dtype=torch.int64, device=dev)
        dist.all_reduce(t_extra, op=dist.ReduceOp.MAX)
        extra = int(t_extra.item())
    for i in range(extra):
        step_dev(warm_eff + i)
    warm_eff += extra
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    fork(e0)
    for i in range(args.steps):
        out = step_dev(i)
    join(e1)
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = world * batch * args.steps / (ms_max / 1e3)

    # ---- end to end through the host-buffer entry (pinned host inputs, H2D + forward + D2H bpp each step) ----
    pin_i = [t.pin_memory() for t in imgs_h[:3]]
    pin_s = [t.pin_memory() for t in scores_h[:3]]
    bpp_pin = [torch.empty(batch, dtype=torch.float32).pin_memory() for _ in range(3)]
    def step_host(i):
        # pinned host inputs -> H2D -> forward -> D2H of the per-image bpp, all on the step's stream
        models[i % S].forward_host(pin_i[i % 3], pin_s[i % 3], bpp_pin[i % 3], stream=side[i % S])

    for i in range(max(args.warmup, S)):
        step_host(i)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fork(e2)
    for i in range(args.steps):
        step_host(i)
    join(e3)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * args.steps / (t.item() / 1e3)
    clocks = sampler.stop() if rank == 0 else None
    h2d = imgs_h[0].numel() * 4 + scores_h[0].numel() * 4
    d2h = batch * 4

    # ---- roofline of the dominant kernel family: CUDA events inside the library --------------------------------
    peaks = load_peaks()
    # events bracket every RUN of consecutive launches of one kernel family (132 of the 133 GEMM-engine launches sit in
    # runs of 2..5: fc1+fc2, g_a, h_a, h_s, the five layers of a cc / lrp net), so the launch-to-launch overlap (PDL) inside
    # a run is kept and the event gap between its members is not charged to the kernel.  The stricter every-launch
    # bracketing is reported next to it as `per_launch_events`.
    def profiled(by_run):
        model.profile(True, by_run=by_run)
        for i in range(2):
            model(imgs_d[i % n_rot], scores_d[i % n_rot])
        torch.cuda.synchronize()
        f_ = model.profile_read()
        model.profile(False)
        return f_
    fams_launch = profiled(False)
    fams = profiled(True)
    tot_ms = sum(f["ms"] for f in fams) or 1.0
    gemm = next((f for f in fams if f["name"] == "gemm_tc"), None)
    roofline = None
    traffic = None
    tr_path = ROOT / "profiles" / f"r01_launches_{wl}.json"      # ncu dram__bytes_read+write per launch (cold cache)
    if tr_path.exists():
        try:
            traffic = json.loads(tr_path.read_text())["families"]["gemm_tc_kernel"]["dram_MB_per_launch"] * 1e6
        except Exception:
            traffic = None
    if gemm and gemm["ms"] > 0:
        achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv engine)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": "bytes per launch, ncu dram__bytes_read+write averaged over the gemm_tc "
                    "launches of one forward (profiles/r01_launches_*.json); algorithmic_flops_per_launch below",
                    "algorithmic_flops_per_launch": gemm["flops"] / gemm["launches"], "peak_source": f"{peaks['source']} bf16_tflops_sustained (of measured)",
                    "launches_per_step": gemm["launches"], "avg_launch_us": gemm["ms"] * 1e3 / gemm["launches"],
                    "share_of_step": gemm["ms"] / tot_ms,
                    "timing": "CUDA events around each run of consecutive gemm_tc launches, inside the library, on the launching stream",
                    "per_launch_events": (lambda g_: {"avg_launch_us": g_["ms"] * 1e3 / g_["launches"],
                                                      "frac": g_["flops"] / (g_["ms"] * 1e-3) / 1e12 / peak})(
                        next(f for f in fams_launch if f["name"] == "gemm_tc")),
                    "families": {f["name"]: {"ms": round(f["ms"], 4), "launches": f["launches"]} for f in fams}}

    # memory-bound kernel families against the measured HBM peak (algorithmic bytes / per-launch event time; these
    # launches move 0.1-19 MB each, i.e. they are launch-latency bound at batch 64 - SURVEY 8d)
    hbm_kernels = {f["name"]: {"GB/s": round(f["bytes"] / (f["ms"] * 1e-3) / 1e9, 1), "launches": f["launches"],
                               "MB_per_launch": round(f["bytes"] / f["launches"] / 1e6, 3),
                               "frac_of_hbm_peak": round(f["bytes"] / (f["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
                   for f in fams_launch if f["bytes"] > 0 and f["ms"] > 0}
    launches = model.launch_count(batch)

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample ---------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_model
        sdc = make_state_dict(cfg, seed=0)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = min(batch, args.ref_sample)
        ref_model.forward_rate(sdc, cfg, imgs_h[0][:sample], scores_h[0][:sample])          # warm-up
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
            ref_model.forward_rate(sdc, cfg, imgs_h[0][:sample], scores_h[0][:sample])
            reps += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sample * reps / dt, "unit": "images/s", "cores": torch.get_num_threads(),
                        "kind": "port", "sample": f"{sample} images x {reps} reps of the same workload, fp32 oracle "
                                                  f"(oracle/ref_model.py) incl. the host python mask routine"}

    if rank == 0:
        line = {
            "metric": "images/s masked-ViT encode+rate", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "name": wl, "per_gpu_batch": batch, "global_batch": batch * world,
                       "parallelism": f"dp{world} (images sharded, weights replicated, 16-byte rate all-reduce)",
                       "streams_per_gpu": S, "warmup_effective": warm_eff,
                       "l2_policy": f"inputs rotate over {n_rot} batches ({n_rot * imgs_h[0].numel() * 4 / 1e6:.0f} MB) "
                                    "+ 350 MB of weights per step > 126 MB L2",
                       "algorithmic_gflop_per_image": ALGO_GFLOP_PER_IMG.get(wl)},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches * args.steps,
            "roofline": roofline, "hbm_kernels": hbm_kernels, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "model_tflops": value * (ALGO_GFLOP_PER_IMG.get(wl) or 0) / 1e3,
            "model_tflops_frac_of_peak": value * (ALGO_GFLOP_PER_IMG.get(wl) or 0) / 1e3 / peaks["bf16_tflops_sustained"],
            "bpp_mean_last_step": out["bpp"].mean().item(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="B64", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=8, help="images per step for the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="batches in flight per GPU (independent handles/streams)")
    ap.add_argument("--batch", type=int, default=0, help="experiment only: images per step instead of the workload's batch")
    args = ap.parse_args()
    kwargs, batch, desc = WORKLOADS[args.workload]
    if args.batch > 0:
        desc = desc.replace(f"batch {batch}", f"batch {args.batch} (EXPERIMENT, not the named workload)")
        batch = args.batch
    if args.impl == "reference":
        run_reference(args, kwargs, batch, desc)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args, kwargs, batch, desc, args.workload)


if __name__ == "__main__":
    main()
