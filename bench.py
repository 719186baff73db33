#!/usr/bin/env python
"""Benchmark of the RDMNet dense-matching hot path (BASELINE.json: scan-pairs/s on the infer.py path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one scan pair through the whole path: GPU pyramid build (4x grid_subsample + 12 radius searches) +
RDMNet.forward (KPConv encoder/decoder, 2x ThDRoFormer, vote + NMS, partition, coarse matching, Sinkhorn, LGR pose).
Workload = BASELINE.json configs[1]: synthetic KITTI pair (~16k points per scan after the 0.3 m downsample, SURVEY.md
8(d) config 2), forward only, fp32. One process per GPU; pairs shard across ranks with no collective on the data path
(weak scaling: every rank runs K steps). Prints ONE JSON line on rank 0.

  value    pairs/s with the raw points already resident in HBM (CUDA events around each step, L2 flushed between
           steps, max over ranks)
  e2e      pairs/s through the host-buffer API (rdmnet_b200.api.PairRegistrar): pinned host points -> H2D -> path ->
           D2H of pose + correspondences, inside the timed region
  roofline KPConv neighbour-gather kernel: algorithmic bytes / CUDA-event time of its launches inside
           the timed steps, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle of the same path (reference C++ core for the pyramid when oracle/_ref was built, torch
           fp32 restatement of the model) on the host cores, bounded sample
--impl reference times that CPU path alone (the reference ships no GPU kernels of its own: SURVEY 2.2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scan_pairs_per_sec_infer_path"
UNIT = "pairs/s"
WORKLOAD = "synthetic KITTI pair (~16k pts/scan after 0.3 m voxel downsample), infer.py path, forward-only"
LIMITS = [65, 63, 69, 70, 81]
CKPT = os.path.join(ROOT, "tests", "golden", "_big", "rdmnet_state.pt")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=4, help="distinct synthetic pairs per rank (cycled)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="one pair at a time (model(data_dict) / PairRegistrar.register) instead of the pair pipeline")
    ap.add_argument("--cpu-pairs", type=int, default=3, help="pairs in the bounded cpu_baseline sample")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE.json configs[4]: 256 synthetic pairs, 64 per size class (4k/8k/16k/32k pts/scan), dealt "
                         "round-robin over the ranks; reports per-class pairs/s and the per-layer gather roofline")
    ap.add_argument("--sweep-pairs", type=int, default=256)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"],
                    help="fp32 (default, the parity mode: 3-term tf32 split) or tf32 (BASELINE config 3 reduced-precision mode: one "
                         "tensor-core product per k-step; NOT the headline number)")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE.json configs[3]: training steps (the reference's unmodified experiments/model.py + loss.py on the "
                         "drop-in, Adam, bucketed bf16 gradient all-reduce over NCCL when N > 1) on synthetic KITTI-shaped pairs")
    ap.add_argument("--train-comm", default="bf16", choices=["bf16", "f32"], help="wire format of the gradient all-reduce")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference's 1-GPU eager measurement")
    return ap.parse_args()


def make_pairs(n, rank):
    from rdmnet_b200 import synthetic
    return [synthetic.make_pair(pair_id=rank * 1000 + i) for i in range(n)]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). Default source: NVML
    in-process (the counters behind `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`), sampled
    synchronously between timed steps (sample_now); BENCH_CLOCKS=smi runs the recipe's nvidia-smi -lms 200 process instead. The external
    process was measured to stall the GPU work of this bench for 50-190 ms per poll often enough to turn one step in
    ~100 into a 10-40x outlier (profiles/README.md), so it is the fallback. Started before the warm-up; mark_begin() /
    stop() select the samples inside the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.samples, self.t0, self.mode, self._stop = index, None, [], None, None, False

    def start(self):
        if os.environ.get("BENCH_CLOCKS", "nvml") != "smi":
            try:
                import pynvml
                pynvml.nvmlInit()
                phys = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[self.index]) if os.environ.get(
                    "CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit() else self.index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self.nv, self.mode = pynvml, "nvml"
                self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
                self.sample_now()
                return
            except Exception:
                self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.t = threading.Thread(target=self._pump_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc, self.mode = None, None

    def sample_now(self):
        """NVML mode: one synchronous sample. bench.py calls it BETWEEN timed steps (every few steps, next to the L2 flush
        and outside the per-step event brackets): a query perturbs the GPU work in flight for milliseconds, and between
        steps the main stream is idle, so the observer effect stays out of the step times."""
        if self.mode != "nvml":
            return
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            self.samples.append((time.time(), sm, self.max_sm, [n for n, bit in names if mask & bit]))
        except Exception:
            pass

    def _pump_smi(self):
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8])
                  if v.lower().startswith("active")]
            self.samples.append((time.time(), sm, mx, rs))

    def mark_begin(self):
        self.t0 = time.time()

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        t1 = time.time()
        if self.mode == "smi":
            time.sleep(0.22)  # let the sample that was being taken at t1 arrive
            self.proc.terminate()
            self.t.join(timeout=2)
        t0 = self.t0 if self.t0 is not None else 0.0
        slack = 0.22 if self.mode == "smi" else 0.0
        inside = [x for x in self.samples if t0 <= x[0] <= t1 + slack]
        if not inside and self.samples:  # region shorter than one polling period: the sample nearest to it
            inside = [min(self.samples, key=lambda x: abs(x[0] - t1))]
        sm = [x[1] for x in inside]
        reasons = sorted({r for x in inside for r in x[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(x[2] for x in inside) if inside else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.mode == "nvml" else "nvidia-smi -lms 200"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def load_state():
    import torch
    if os.path.exists(CKPT):
        return torch.load(CKPT, map_location="cpu", weights_only=True), "pretrained reference checkpoint"
    from rdmnet_b200.model import create_model
    torch.manual_seed(7351)
    return {k: v.detach().clone() for k, v in create_model().state_dict().items()}, "random-init weights (seed 7351)"


def cpu_reference_pair(state, pair, impl):
    """The reference's CPU path for one pair: registration_collate_fn_stack_mode (C++ ext) + RDMNet.forward (torch CPU),
    restated by oracle/ (checker code; this is the one place outside tests/ allowed to execute it)."""
    import torch
    from oracle import model_oracle as MO
    from oracle import pyramid as OP
    pts = np.concatenate([pair["ref_points"], pair["src_points"]])
    lens = [len(pair["ref_points"]), len(pair["src_points"])]
    pyr = OP.precompute_pyramid(pts, lens, 5, 0.3, 4.25 * 0.3, LIMITS, impl)
    tp = MO.pyramid_to_torch(pyr)
    with torch.no_grad():
        out = MO.forward(state, tp, lambda p, l: OP.radius_search(p.numpy(), p.numpy(), l.numpy(), l.numpy(), 2.4,
                                                                  LIMITS[-1], impl))
    return out["estimated_transform"]


def cpu_arm(pairs, n_steps, n_warm, budget_s):
    """-> (cpu_baseline dict, pairs done, seconds, info). Runs the reference's OWN code (baseline/_ref/RDMNet: its collate over
    its C++ extension core + experiments/model_infer.RDMNet on torch CPU) when it is staged - kind "reference" -, else the
    oracle port - kind "port"."""
    import torch
    from oracle import pyramid as OP
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    state, wdesc = load_state()
    info = {}
    try:
        from baseline import reference_runner as RR
        ok, why = RR.available()
    except Exception as e:  # pragma: no cover
        ok, why = False, repr(e)
    if ok:
        model, cfg = RR.build_model(state, LIMITS, "cpu")
        done, dt, T, n0 = RR.run_cpu(model, cfg, pairs, n_steps, n_warm, budget_s)
        info = {"last_pair": (n_warm + done - 1) % len(pairs), "estimated_transform": T}
        kind = "reference"
        sample = (f"{done} pair(s) of the workload after {n_warm} warm-up through the UNMODIFIED reference (baseline/_ref/RDMNet): "
                  f"registration_collate_fn_stack_mode over its C++ extension core (1 thread, as in a DataLoader worker) + "
                  f"experiments/model_infer.RDMNet.forward on torch CPU with {cores} threads; {wdesc}")
    else:
        impl = "ref" if OP.ref_available() else "port"
        for i in range(n_warm):
            cpu_reference_pair(state, pairs[i % len(pairs)], impl)
        t0, done, T = time.perf_counter(), 0, None
        for i in range(n_warm, n_warm + n_steps):
            T = cpu_reference_pair(state, pairs[i % len(pairs)], impl)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        info = {"last_pair": (n_warm + done - 1) % len(pairs), "estimated_transform": T.numpy()}
        kind = "port"
        sample = (f"{done} pair(s) of the workload after {n_warm} warm-up ({why}); pyramid = "
                  f"{'reference C++ core (oracle/_ref), 1 thread' if impl == 'ref' else 'C restatement'}"
                  f"; model forward = torch-CPU fp32 restatement of experiments/model_infer.py on {cores} threads; {wdesc}")
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, done, dt, info


def reference_gpu_eager(pairs, n_steps=48, n_warm=8, workers=8):
    """BASELINE.md 3.4 / the north star's denominator: the unmodified reference, PyTorch-eager on ONE GPU, fed by 8 CPU
    collate workers. None when the reference tree is not staged or there is no GPU."""
    import torch
    if not torch.cuda.is_available():
        return None
    try:
        from baseline import reference_runner as RR
        ok, why = RR.available()
        if not ok:
            return {"unavailable": why}
        state, wdesc = load_state()
        model, cfg = RR.build_model(state, LIMITS, "cuda")
        done, dt, fwd_ms, T = RR.run_gpu_eager(model, cfg, pairs, n_steps, n_warm, workers)
        del model
        torch.cuda.empty_cache()
        return {"value": done / dt, "unit": UNIT, "forward_ms_cuda_events": fwd_ms, "pairs": done, "warmup": n_warm,
                "collate_workers": workers,
                "what": "unmodified experiments/model_infer.RDMNet, PyTorch eager fp32 on 1 GPU, CPU collate (its C++ ext core) in "
                        f"{workers} DataLoader workers, wall clock over the timed pairs (BASELINE.md 3.4); {wdesc}"}
    except Exception as e:
        return {"unavailable": repr(e)[:300]}


def run_reference_train(args):
    """`--train --impl reference`: the unmodified reference's training step on one GPU (baseline/reference_runner.run_gpu_train)."""
    import torch
    from baseline import reference_runner as RR
    ok, why = RR.available()
    if not ok or not torch.cuda.is_available():
        emit(json.dumps({"impl": "reference", "unavailable": why or "no CUDA device"}))
        return
    pairs = make_pairs(args.pairs, 0)
    state, wdesc = load_state()
    done, dt, dev_ms, (l0, l1) = RR.run_gpu_train(state, LIMITS, pairs, args.steps, max(args.warmup, 3))
    emit(json.dumps({"impl": "reference", "metric": "train_pairs_per_sec", "value": done / dt, "unit": "pairs/s", "n_gpus": 1, "steps": done,
                     "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                     "config": {"workload": "config 4: training step on a synthetic KITTI pair (~16k pts/scan), batch 1, the UNMODIFIED "
                                            "reference: experiments/model.py + loss.py + Adam, PyTorch eager on one GPU, CPU collate in 8 "
                                            "DataLoader workers", "weights": wdesc + " as the starting point", "neighbor_limits": LIMITS},
                     "device_ms_per_step": dev_ms, "loss_first_last": [l0, l1],
                     "e2e": {"value": done / dt, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.train:
        run_reference_train(args)
        return
    pairs = make_pairs(args.pairs, 0)  # the same pairs, step count and warm-up as the repo arm
    cb, done, dt, _ = cpu_arm(pairs, args.steps, args.warmup, budget_s=240.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": 1, "distinct_pairs_per_rank": len(pairs),
                       "points_per_pair": [int(len(p["ref_points"]) + len(p["src_points"])) for p in pairs],
                       "neighbor_limits": LIMITS,
                       "note": "CPU path of the reference (it ships no GPU kernels); rank 0 only"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_reference_gpu:
        line["reference_gpu_eager"] = reference_gpu_eager(pairs)
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from rdmnet_b200 import _lib as L
    from rdmnet_b200.api import PairRegistrar, PairStreamRegistrar
    from rdmnet_b200.model import PairPipeline
    from rdmnet_b200.ops import kpconv_gather_bytes as ops_kpconv_gather_bytes
    from rdmnet_b200.model import create_model

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: rdmnet_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.lib()  # fail loudly if librdm_sm100.so is missing
    import rdmnet_b200
    rdmnet_b200.set_precision(args.precision)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pairs = make_pairs(args.pairs, rank)
    model = create_model()
    if os.path.exists(CKPT):
        model.load_state_dict(torch.load(CKPT, map_location="cpu", weights_only=True), strict=True)
        wdesc = "pretrained reference checkpoint"
    else:
        torch.manual_seed(7351)
        wdesc = "random-init weights (seed 7351)"
    model = model.to(dev).eval()

    d_pairs = []
    for p in pairs:
        pts = torch.from_numpy(np.concatenate([p["ref_points"], p["src_points"]])).to(dev)
        lens = torch.tensor([len(p["ref_points"]), len(p["src_points"])], dtype=torch.int64, device=dev)
        d_pairs.append((pts, lens))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    pipelined = not args.no_pipeline
    pipe = PairPipeline(model, dev)

    def step(i):
        pts, lens = d_pairs[i % len(d_pairs)]
        return model({"points": pts, "lengths": lens})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()  # before the warm-up: see ClockSampler
    # the warm-up runs exactly what the timed steps run (flush kernel, profiler brackets, clock query): their first
    # use loads kernels / grows pools, which showed up as 10-17 ms outliers on timed steps 1-2
    prof_mask = 0 if os.environ.get("BENCH_NO_PROF", "0") == "1" else int(os.environ.get("BENCH_PROF_MASK", "1"))
    L.prof_enable(prof_mask)

    def warm_hook(i):
        clocks.sample_now()
        flush.fill_(i & 0xFF)

    # set-up, not warm-up: every distinct pair once more than the allocator needs to have a cached block for each of its
    # size classes in the pipelined interleaving (a first-time cudaMalloc inside the timed region is a 15-35 ms step)
    n_prime = 2 * len(d_pairs) + 1
    n_lead = n_prime + args.warmup
    gen_main = None
    if pipelined and not pipe.overlap:
        # ONE generator for set-up, warm-up and the timed steps: the pipeline is not torn down and refilled at the start of the
        # timed region (a refill puts the first pair's pyramid build, serially, into timed step 0). The barrier + synchronize
        # between warm-up and timed steps happens with the look-ahead pyramid of the first timed pair already built - as in
        # every later step, whose pyramid was built during its predecessor.
        timed_before = [None]

        def lead_before(i):
            if i < n_lead:
                warm_hook(i)
            elif timed_before[0] is not None:
                timed_before[0](i - n_lead)
        gen_main = pipe.run((d_pairs[i % len(d_pairs)] for i in range(n_lead + args.steps + 1)), before_step=lead_before)
        for _ in range(n_lead):
            next(gen_main)
    elif pipelined:
        for _ in pipe.run((d_pairs[i % len(d_pairs)] for i in range(n_prime + args.warmup)), before_step=warm_hook):
            pass
    else:
        for i in range(n_prime + args.warmup):
            warm_hook(i)
            step(i)
    barrier()

    # ---- timed region 1: device-resident inputs; per-step CUDA events, L2 flush (untimed) between steps.
    # Pipelined: while step i is in the network the pyramid of step i+1 is built on a side stream, so every timed step
    # contains exactly one pyramid build and one network pass; the events live on the main stream, which also waits
    # (inside the bracket) for the pyramid it consumes.
    # in-library CUDA-event brackets around the KPConv gather launches only (mask 1): every bracket is two stream operations
    # that also interrupt the programmatic-dependent-launch chain, so the informational weight-GEMM brackets stay off
    L.prof_enable(prof_mask)  # (re-enabling drops the warm-up's records)
    no_flush = os.environ.get("BENCH_NO_FLUSH", "0") == "1"    # debug knob (the reported configuration always flushes)
    launches0 = L.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:  # torch creates the CUDA event at its first record(): do that here, not inside the timed loop
        a.record()
        b.record()
    barrier()
    # the cyclic garbage collector stays out of the timed regions: its generation-2 passes (count-triggered, so they hit
    # the same step indices run after run) were 15-80 ms pauses of the single host thread that drives the GPU
    import gc
    gc.collect()
    gc.disable()
    clocks.mark_begin()
    if pipelined and pipe.overlap:
        clocks.sample_now()  # edge sample (see `before` below)
    t_wall0 = time.perf_counter()
    if pipelined and pipe.overlap:
        # (opt-in RDM_PIPE_OVERLAP=1) K pairs through the pair pipeline, from an idle GPU to an idle GPU (pipeline fill and drain included). Three pairs are in
        # flight (pyramid of i+2, backbone of i+1, matching tail of i), so steps overlap: the time of the K steps is taken between
        # one event before the first pair enters and the event after the last pair's last kernel, and the L2 flush between pairs is
        # queued on the entering pair's stream INSIDE the timed region (it overlaps the previous pair's tail).
        t_first = torch.cuda.Event(enable_timing=True)
        t_first.record()

        def before(i):
            # ONE clock / throttle query under load, in the middle of the region: with three pairs in flight there is no idle moment
            # to hide an NVML query in, and each one stalls the launch stream for 5-40 ms (steps 3 and 9 were the outliers of every
            # run when the region was sampled every 6 steps); two more samples sit right at the region's edges (below)
            if i == args.steps // 2:
                clocks.sample_now()
            if not no_flush:
                flush.fill_(i & 0xFF)
            ev[i][0].record()

        def after(i):
            ev[i][1].record()
        gen = pipe.run((d_pairs[i % len(d_pairs)] for i in range(args.steps)), before_step=before, after_step=after)
        for i in range(args.steps):
            out = next(gen)
        for _ in gen:
            pass
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks.sample_now()
        launches = L.launch_count() - launches0
    elif pipelined:
        def before(i):
            if i < args.steps:
                if i % 6 == 3:
                    clocks.sample_now()  # under load (the side stream is building the next pyramid), outside the brackets
                if not no_flush:
                    flush.fill_(i & 0xFF)
                ev[i][0].record()
        timed_before[0] = before
        gen = gen_main
        for i in range(args.steps):
            out = next(gen)
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = L.launch_count() - launches0
        for _ in gen:  # the look-ahead pair: untimed
            pass
    else:
        for i in range(args.steps):
            if i % 6 == 3:
                clocks.sample_now()
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            out = step(i)
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = L.launch_count() - launches0
    gc.enable()
    prof = L.prof_read()
    L.prof_enable(False)
    torch.cuda.synchronize()
    if pipelined and pipe.overlap:
        ends = [t_first] + [b for _, b in ev]
        step_ms = [ends[i].elapsed_time(ends[i + 1]) for i in range(args.steps)]  # completion-to-completion intervals
        total_ms = float(t_first.elapsed_time(ev[-1][1]))
    else:
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(sum(step_ms))
    clk = clocks.stop()

    # ---- timed region 2: end to end through the host-buffer API (pinned H2D of the points, D2H of the results)
    maxp = max(p[0].shape[0] for p in d_pairs)
    host_pairs = [(pairs[i % len(pairs)]["ref_points"], pairs[i % len(pairs)]["src_points"]) for i in range(max(args.steps, 3))]
    if pipelined:
        reg = PairStreamRegistrar(model, max_points=maxp, device=dev)
        for res in reg.register_stream([host_pairs[i % len(host_pairs)] for i in range(n_lead)]):
            pass  # warm-up: every distinct pair through the host-buffer path (pinned staging + allocator blocks of its sizes)
        barrier()
        gc.collect()
        gc.disable()
        e2e_allocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        t0 = time.perf_counter()
        e2e_marks = [t0]
        for res in reg.register_stream(host_pairs[:args.steps]):
            e2e_marks.append(time.perf_counter())
        barrier()
        e2e_s = time.perf_counter() - t0
        gc.enable()
        api_name = "rdmnet_b200.api.PairStreamRegistrar.register_stream (host numpy in, host numpy out, pair i+1 staged while pair i runs)"
    else:
        reg = PairRegistrar(model, max_points=maxp, device=dev)
        for i in range(min(args.warmup, 3)):
            reg.register(*host_pairs[i])
        barrier()
        gc.collect()
        gc.disable()
        t0 = time.perf_counter()
        for i in range(args.steps):
            res = reg.register(*host_pairs[i])
        barrier()
        e2e_s = time.perf_counter() - t0
        gc.enable()
        api_name = "rdmnet_b200.api.PairRegistrar.register (host numpy in, host numpy out)"

    e2e_intervals = None
    if pipelined:
        iv = np.diff(np.asarray(e2e_marks)) * 1e3
        e2e_intervals = {"cudaMallocs": int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - e2e_allocs0),
                         "min": float(iv.min()), "median": float(np.median(iv)), "max": float(iv.max()),
                         "over_2x_median": [[int(i), round(float(t), 1)] for i, t in enumerate(iv) if t > 2 * np.median(iv)]}
    # ---- untimed extra pass (rank 0): the KPConv weight GEMMs bracketed with CUDA events for the tensor-pipe roofline entry
    # (the brackets break the programmatic-dependent-launch chain, which is why they are off inside the timed regions),
    # and one plain forward per distinct pair whose pose / correspondence count the cpu_baseline leg is compared with
    gemm_prof, my_poses = [], {}
    if rank == 0 and not args.no_cpu_baseline:
        L.prof_enable(2)
        for i in range(len(d_pairs)):
            o = step(i)
            my_poses[i] = (o["estimated_transform"].cpu().numpy(), int(o["corr_scores"].shape[0]))
        torch.cuda.synchronize()
        gemm_prof = [r for r in L.prof_read() if r[0] == 2]
        L.prof_enable(False)

    # max over ranks
    tm = torch.tensor([total_ms, e2e_s * 1e3, t_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, wall_ms = tm.tolist()

    # registration quality on the last pair against the synthetic ground truth (the reference's success criterion is
    # RRE < 5 deg and RTE < 2 m, experiments/config.py:66-67)
    def pose_err(T, Tg):
        T, Tg = np.asarray(T, np.float64), np.asarray(Tg, np.float64)
        return (float(np.degrees(np.arccos(np.clip((np.trace(T[:3, :3].T @ Tg[:3, :3]) - 1) / 2, -1, 1)))),
                float(np.linalg.norm(T[:3, 3] - Tg[:3, 3])))
    rre, rte = pose_err(res["estimated_transform"], pairs[(args.steps - 1) % len(pairs)]["transform"])

    if rank == 0:
        peak, peak_src = peaks()
        # KPConv of RDMNet: C_out == C_in except encoder1_1 (1 -> 64) (experiments/backbone.py:11-70)
        g = {"launches": 0, "ms": 0.0, "bytes": 0}
        w = {"launches": 0, "ms": 0.0}
        per_layer = {}
        for tag, ms, m_, n_, h_, c_ in prof:
            if tag == 1:
                nb = ops_kpconv_gather_bytes(m_, h_, c_, 64 if c_ == 1 else c_, 4)
                g["launches"] += 1; g["ms"] += ms; g["bytes"] += nb
                d = per_layer.setdefault("M%d_H%d_C%d" % (m_, h_, c_), [0.0, 0, 0])
                d[0] += ms; d[1] += nb; d[2] += 1
            elif tag == 2:
                w["launches"] += 1; w["ms"] += ms
        ach = (g["bytes"] / 1e9) / (g["ms"] / 1e3) if g["ms"] > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "kpconv_gather_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * args.steps / (total_ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "tf32 contractions, f32 elsewhere (reduced-precision mode, config 3)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": 1, "distinct_pairs_per_rank": len(pairs), "precision": args.precision,
                       "points_per_pair": [int(p[0].shape[0]) for p in d_pairs], "neighbor_limits": LIMITS,
                       "weights": wdesc,
                       "l2": ("256 MiB flush write queued before every pair enters the network, INSIDE the timed region (pairs overlap)"
                              if (pipelined and pipe.overlap) else "256 MiB flush write between timed steps (untimed)"),
                       "timing": ("K pairs through the pair pipeline from an idle GPU to an idle GPU (fill and drain included): CUDA events "
                                  "before the first pair enters and after the last pair's last kernel; max over ranks; kernel events "
                                  "recorded inside the library around the gather launches" if (pipelined and pipe.overlap) else
                                  "CUDA events per step on the launch stream, summed; max over ranks; kernel events recorded inside the "
                                  "library around the launches"),
                       "sharding": "pairs round-robin over ranks, no data-path collective",
                       "setup_steps_before_warmup": n_prime,
                       "pipeline": (("pair pipeline, three pairs in flight: pyramid of pair i+2 (side stream), backbone of pair i+1 and matching "
                                     "tail of pair i (two network streams); the tail and the searches wait for the encoder of pair i+1, so "
                                     "the KPConv gathers run alone" if pipe.overlap else
                                     "pair pipeline: pyramid of pair i+1 on a side stream during the network pass of pair i")
                                    if pipelined else "off: one pair at a time")},
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
            "step_ms_stats": {"min": float(np.min(step_ms)), "median": float(np.median(step_ms)), "max": float(np.max(step_ms)),
                              "outlier_steps": [int(i) for i, t in enumerate(step_ms) if t > 1.5 * float(np.median(step_ms))]},
            "e2e": {"value": world * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": reg.h2d_bytes,
                    "d2h_bytes_per_step": reg.d2h_bytes, "ms_per_step": e2e_ms / args.steps,
                    "result_interval_ms": e2e_intervals,
                    "api": api_name},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "kpconv_gather_sparse_kernel (13 launches/step) + kpconv_gather_c1_kernel (1)", "bound": "hbm",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "peak_source": peak_src, "launches": g["launches"],
                         "avg_launch_us": 1e3 * g["ms"] / max(g["launches"], 1),
                         "algorithmic_bytes_per_launch": g["bytes"] / max(g["launches"], 1),
                         "definition": "G = M*H*(C_in*4+12+4) + M*(C_out*4+12) per launch (SURVEY 8(d), fp32, int32 idx)",
                         # the look-ahead pair's gathers are bracketed too: normalise by the launches recorded, not by K
                         "kpconv_gather_ms_per_step": g["ms"] / max(g["launches"] / 14.0, 1e-9),
                         "per_layer_GBps": {k: (v[1] / 1e9) / (v[0] / 1e3) for k, v in per_layer.items() if v[0] > 0}},
            "clocks": clk,
            "pose_vs_synthetic_gt": {"rre_deg": rre, "rte_m": rte, "n_corr": int(res["corr_scores"].shape[0]),
                                     "registered": bool(rre < 5.0 and rte < 2.0)},
        }
        if gemm_prof:
            # tensor-pipe roofline of the KPConv weight contraction (M x 15 C_in) . (15 C_in x C_out), tcgen05 kind::tf32 with
            # the 3-term split: `achieved` counts the useful fp32-equivalent flops 2 M K N; the tensor pipe issues 3x that in
            # tf32 MMAs. Peak: MEASURED_PEAKS.json has no tf32 figure; dense tf32 is half the bf16 rate on this part
            # (B200_PROFILING.md: 1.1 vs 2.25 PFLOP/s nominal), so peak = bf16_tflops_sustained / 2 (assumption stated).
            fl = sum(2.0 * m_ * n_ * c_ for _, _, m_, n_, _, c_ in gemm_prof)
            ms = sum(r[1] for r in gemm_prof)
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    bf16 = float(json.load(f)["bf16_tflops_sustained"])
                psrc = "measured bf16_tflops_sustained / 2 (tf32 dense = half the bf16 rate: assumption)"
            except Exception:
                bf16, psrc = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 / 2"
            ach = fl / 1e12 / (ms / 1e3)
            passes = 3.0 if args.precision == "fp32" else 1.0
            kname = ("gemm_tf32x3 (tcgen05 kind::tf32, 3-term split, A in TMEM)" if args.precision == "fp32"
                     else "gemm_tf32x1 (tcgen05 kind::tf32, single product, no operand split)")
            line["roofline_gemm"] = {"kernel": kname + ", KPConv weight contraction, 14 launches/step",
                                     "bound": "tensor", "achieved": ach, "issued_tf32_tflops": passes * ach, "peak": bf16 / 2.0,
                                     "unit": "TFLOP/s", "frac": ach / (bf16 / 2.0), "frac_issued": passes * ach / (bf16 / 2.0),
                                     "peak_source": psrc, "launches": len(gemm_prof),
                                     "kpconv_weight_gemm_ms_per_step": ms / max(len(gemm_prof) / 14.0, 1e-9),
                                     "note": "measured in an untimed extra pass with event brackets around each launch"}
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _, info = cpu_arm(pairs, args.cpu_pairs, 1, budget_s=60.0)
            line["cpu_baseline"] = cb
            # parity on the BENCHED pairs: this arm's pose vs the CPU reference's on the same pair
            if info.get("estimated_transform") is not None and info["last_pair"] in my_poses:
                Tm, nc = my_poses[info["last_pair"]]
                Tc = np.asarray(info["estimated_transform"], np.float64)
                d_rre, d_rte = pose_err(Tm, Tc)
                line["parity_vs_cpu"] = {"pair": int(info["last_pair"]), "pose_max_abs_err": float(np.abs(Tm - Tc).max()),
                                         "pose_rel_err": float(np.abs(Tm - Tc).max() / np.abs(Tc).max()),
                                         "rre_between_deg": d_rre, "rte_between_m": d_rte, "n_corr_gpu": nc,
                                         "cpu_kind": cb["kind"]}
            if not args.no_reference_gpu:
                rg = reference_gpu_eager(pairs)
                line["reference_gpu_eager"] = rg
                if rg and "value" in rg:
                    line["vs_reference_gpu_eager"] = {"e2e_ratio": line["e2e"]["value"] / rg["value"],
                                                      "value_ratio": line["value"] / rg["value"],
                                                      "target": ">= 10x (BASELINE.json north_star)"}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- config 5: size sweep
def run_sweep(args, rank, world, local_rank):
    """BASELINE.json configs[4] / SURVEY 8(d) config 5: `--sweep-pairs` synthetic pairs, a quarter per size class (4k / 8k /
    16k / 32k points per scan), pair ids dealt round-robin over the ranks (rdmnet_b200.sharding.shard_pair_ids, no
    collective on the data path). Per class: pairs/s (device time: CUDA events per pair inside the pair pipeline, L2 flushed
    between pairs; whole job = pairs of all ranks / slowest rank) and the KPConv gather roofline on G."""
    import torch
    import torch.distributed as dist
    from rdmnet_b200 import _lib as L
    from rdmnet_b200 import synthetic
    from rdmnet_b200.model import PairPipeline, create_model
    from rdmnet_b200.ops import kpconv_gather_bytes
    from rdmnet_b200.sharding import shard_pair_ids
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: rdmnet_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.lib()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = create_model()
    if os.path.exists(CKPT):
        model.load_state_dict(torch.load(CKPT, map_location="cpu", weights_only=True), strict=True)
    model = model.to(dev).eval()
    classes = list(synthetic.SIZE_CLASSES)
    per_class = args.sweep_pairs // len(classes)
    distinct = 4  # distinct scene geometries per class (ray casting a scene costs seconds of host time); cycled
    mine = shard_pair_ids(args.sweep_pairs, rank, world)  # pair id p: class p // per_class
    geo = {}
    for c in classes:
        ne, na = synthetic.SIZE_CLASSES[c]
        for g_ in range(distinct):
            pr = synthetic.make_pair(pair_id=5000 + g_, n_elev=ne, n_azim=na)
            pts = torch.from_numpy(np.concatenate([pr["ref_points"], pr["src_points"]])).to(dev)
            lens = torch.tensor([len(pr["ref_points"]), len(pr["src_points"])], dtype=torch.int64, device=dev)
            geo[(c, g_)] = (pts, lens)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pipe = PairPipeline(model, dev)
    peak, peak_src = peaks()
    res = {}
    import gc
    for ci, c in enumerate(classes):
        ids = [p_ for p_ in mine if p_ // per_class == ci]
        items = [geo[(c, p_ % distinct)] for p_ in ids]
        for _ in pipe.run(items[:3] + [geo[(c, g_)] for g_ in range(distinct)] * 2, before_step=lambda i: flush.fill_(i & 0xFF)):
            pass  # warm-up + allocator priming for this size class
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        L.prof_enable(1)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in items]
        for a, b in ev:
            a.record(); b.record()
        torch.cuda.synchronize()
        gc.collect(); gc.disable()

        n_corr = 0
        allocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        if pipe.overlap:
            t_first = torch.cuda.Event(enable_timing=True)
            t_first.record()

            def before(i):
                flush.fill_(i & 0xFF)

            def after(i, ev=ev):
                ev[i][1].record()
            for out in pipe.run(items, before_step=before, after_step=after):
                n_corr += int(out["corr_scores"].shape[0])
            torch.cuda.synchronize()
            ms = float(t_first.elapsed_time(ev[-1][1]))  # all pairs of the class through the pipeline, fill and drain included
        else:
            def before(i, ev=ev):
                if i < len(ev):
                    flush.fill_(i & 0xFF)
                    ev[i][0].record()
            gen = pipe.run(items + items[:1], before_step=before)
            for i in range(len(items)):
                out = next(gen)
                ev[i][1].record()
                n_corr += int(out["corr_scores"].shape[0])
            torch.cuda.synchronize()
            for _ in gen:
                pass
            ms = float(sum(a.elapsed_time(b) for a, b in ev))
        gc.enable()
        prof = [r for r in L.prof_read() if r[0] == 1]
        L.prof_enable(False)
        gb = sum(kpconv_gather_bytes(m_, h_, c_, 64 if c_ == 1 else c_, 4) for _, _, m_, _, h_, c_ in prof)
        gms = sum(r[1] for r in prof)
        t = torch.tensor([float(len(items)), ms, float(gb), gms, float(n_corr), float(sum(int(x[0].shape[0]) for x in items))],
                         dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            allt = torch.stack(allt).cpu().numpy()
        else:
            allt = t.cpu().numpy()[None]
        n_tot, t_max = float(allt[:, 0].sum()), float(allt[:, 1].max())
        res[c] = {"pairs": int(n_tot), "pairs_per_s": n_tot / (t_max / 1e3), "ms_per_pair_per_gpu": float(allt[:, 1].sum() / n_tot),
                  "mean_points_per_pair": float(allt[:, 5].sum() / n_tot), "mean_corr": float(allt[:, 4].sum() / n_tot),
                  "gather_GBps": float(allt[:, 2].sum() / 1e9 / (allt[:, 3].sum() / 1e3)),
                  "gather_frac": float(allt[:, 2].sum() / 1e9 / (allt[:, 3].sum() / 1e3) / peak),
                  "per_rank_ms": [float(x) for x in allt[:, 1]],
                  "cudaMallocs_in_timed_region_rank0": int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - allocs0)}
    if rank == 0:
        n_all = sum(v["pairs"] for v in res.values())
        t_all = sum(max(v["per_rank_ms"]) for v in res.values())
        line = {"metric": METRIC, "value": n_all / (t_all / 1e3), "unit": UNIT, "n_gpus": world, "steps": n_all, "warmup": 3 + 2 * distinct,
                "ms_per_step": t_all / max(n_all / world, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "config 5: %d-pair synthetic sweep, %d pairs per size class %s (points/scan), round-robin over ranks"
                                       % (args.sweep_pairs, per_class, classes),
                           "distinct_geometries_per_class": distinct, "neighbor_limits": LIMITS,
                           "l2": "256 MiB flush write between pairs (untimed)",
                           "timing": "CUDA events per pair on the launch stream; per class: all pairs / slowest rank"},
                "sweep": res, "roofline_peak": {"hbm_gbs": peak, "source": peak_src}}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- config 4: training
def run_train(args, rank, world, local_rank):
    """BASELINE.json configs[3] / SURVEY 8(f).1: one step = one synthetic KITTI-shaped pair per GPU through
    host points -> GPU collate (pyramid) -> experiments/model.py forward in train mode (ground-truth node correspondences, 128
    sampled patches) -> experiments/loss.py OverallLoss -> backward through the kernels of csrc/backward.cu + train.cu ->
    bucketed gradient all-reduce (rdmnet_b200.ddp, overlapped with the backward pass) -> Adam (experiments/trainval.py:34).
    model.py and loss.py are the reference's files, unmodified, bound to rdmnet_b200 by rdmnet_b200.dropin. Weak scaling: every
    rank trains on its own pairs; value = pairs of all ranks / slowest rank's time."""
    import importlib
    import torch
    import torch.distributed as dist
    from rdmnet_b200 import _lib as L
    from rdmnet_b200 import ddp, dropin
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: rdmnet_b200 has no CPU path")
    ref_root = os.path.join(ROOT, "baseline", "_ref", "RDMNet")
    if not os.path.isdir(os.path.join(ref_root, "experiments")):
        raise RuntimeError("--train runs the reference's experiments/model.py + loss.py: stage them with __graft_entry__.build()")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.lib()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dropin.install(reference_root=ref_root)
    config, model_mod, loss_mod = (importlib.import_module(m) for m in ("config", "model", "loss"))
    data = importlib.import_module("geotransformer.utils.data")
    cfg = config.make_cfg()
    cfg.test.vis = False
    cfg.neighbor_limits = LIMITS
    model = model_mod.create_model(cfg)
    if os.path.exists(CKPT):  # a trained starting point: NMS survivors / correspondence counts are those of real training steps
        model.load_state_dict(torch.load(CKPT, map_location="cpu", weights_only=True), strict=True)
        wdesc = "pretrained reference checkpoint as the starting point"
    else:
        wdesc = "random init"
    model = model.to(dev).train()
    loss_fn = loss_mod.OverallLoss(cfg).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-6, fused=True)  # trainval.py:34 (fused = same update, one launch)
    comm_dtype = torch.bfloat16 if args.train_comm == "bf16" else torch.float32
    red = ddp.BucketedGradAllReduce(model, bucket_bytes=8 << 20, comm_dtype=comm_dtype) if world > 1 else None
    pairs = make_pairs(args.pairs, rank)
    items = [dict(ref_points=p["ref_points"], src_points=p["src_points"], ref_feats=np.ones((len(p["ref_points"]), 1), np.float32),
                  src_feats=np.ones((len(p["src_points"]), 1), np.float32), transform=p["transform"].astype(np.float32)) for p in pairs]
    names = ("collate", "forward", "loss", "backward", "allreduce_exposed", "optimizer")
    launches0 = None
    h2d = d2h = 0

    def step(i, ev=None):
        nonlocal h2d, d2h
        it = items[i % len(items)]
        mark = (lambda k: ev[k].record()) if ev is not None else (lambda k: None)
        mark(0)
        dd = data.registration_collate_fn_stack_mode([it], 5, 0.3, 4.25 * 0.3, LIMITS)
        dd = {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v))
              for k, v in dd.items()}
        dd["testing"] = False
        mark(1)
        np.random.seed(1000 + i)
        out = model(dd)
        mark(2)
        with torch.device(dev):  # loss.py:240-243 builds index helpers with bare torch.arange (torch 1.8 tolerated the device mix)
            losses = loss_fn(out, dd)
        mark(3)
        if red is not None:
            red.zero_grad()  # gradients are views into the reducer's flat buckets: zero in place
        else:
            opt.zero_grad(set_to_none=True)
        losses["loss"].backward()
        mark(4)
        if red is not None:
            red.finish()
        mark(5)
        opt.step()
        mark(6)
        h2d = sum(int(v.nbytes) for v in it.values())
        d2h = 4
        return float(losses["loss"].detach())  # the D2H read of the step's result (the reference logs it every iteration)

    for i in range(max(args.warmup, 3)):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = L.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark_begin()
    t0 = time.perf_counter()
    loss_vals = []
    for i in range(args.steps):
        loss_vals.append(step(args.warmup + i, evs[i]))
        if i % 8 == 7:
            sampler.sample_now()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = L.launch_count() - launches0
    phase = np.array([[e[k].elapsed_time(e[k + 1]) for k in range(6)] for e in evs])  # ms
    # the collectives alone (ranks aligned by a barrier, nothing else in flight): separates wire time from rank skew in `exposed`
    alone_ms = 0.0
    if red is not None:
        reps = 5
        dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            for wbuf in red.wire:
                dist.all_reduce(wbuf)
        a1.record()
        torch.cuda.synchronize()
        alone_ms = a0.elapsed_time(a1) / reps
    if rank == 0 and os.environ.get("BENCH_TRAIN_PROFILE"):  # debugging aid: kernel-time table of two more (untimed) steps -> stderr
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(2):
                step(args.warmup + args.steps + i)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70), file=sys.stderr)
    dev_ms = float(phase.sum())
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    if rank == 0:
        per = {n: float(phase[:, k].mean()) for k, n in enumerate(names)}
        line = {"metric": "train_pairs_per_sec", "value": world * args.steps / (wall_ms_max / 1e3), "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": wall_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (gradients on the wire: %s)" % args.train_comm, "data": "synthetic",
                "config": {"workload": "config 4: training step on a synthetic KITTI pair (~16k pts/scan), batch 1 per GPU, "
                                       "experiments/model.py + loss.py unmodified on rdmnet_b200.dropin, Adam lr 1e-4",
                           "parallelism": "dp%d" % world, "distinct_pairs_per_rank": len(items), "weights": wdesc,
                           "timing": "wall clock over the timed steps incl. H2D of the points and the D2H read of the loss, max over ranks; "
                                     "phases = CUDA events on the launch stream (rank 0)",
                           "neighbor_limits": LIMITS},
                "device_ms_per_step": dev_ms_max / args.steps, "phases_ms": per,
                "e2e": {"value": world * args.steps / (wall_ms_max / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "allreduce": {"bytes_per_step": red.bytes_per_step if red is not None else 0, "buckets": len(red.buckets) if red is not None else 0,
                              "wire_dtype": args.train_comm, "exposed_ms": per["allreduce_exposed"], "collectives_alone_ms": alone_ms,
                              "exposed_frac_of_step": per["allreduce_exposed"] / max(dev_ms / args.steps, 1e-9),
                              "note": "exposed = launch-stream time between the end of backward and the averaged gradients being in place: the "
                                      "wire time of the last bucket PLUS the wait for the slower rank (ranks train on different pairs and "
                                      "are host-bound); collectives_alone_ms = all buckets back to back after a barrier"},
                "loss_first_last": [loss_vals[0], loss_vals[-1]], "clocks": clocks}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process' original stdout; everything else that libraries print to fd 1 (NCCL's
    version banner under NCCL_DEBUG=VERSION, for instance) was redirected to stderr in main()."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.train:
        run_train(args, rank, world, local_rank)
    elif args.sweep:
        run_sweep(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
