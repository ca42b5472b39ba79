#!/usr/bin/env python
"""bench.py -- MSDeformAttn fwd+bwd throughput on B200 (queries/s and HBM GB/s), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--loc-dist D]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (default ``detr_encoder_800x1333``, BASELINE.json configs[2], the shape the metric's target is quoted on):
  per GPU N=16 images, S=Lq=22223 multi-scale tokens (levels 100x167, 50x84, 25x42, 13x21), M=8 heads, D=32,
  L=P=4, fp32, and one "step" = 6 independent layer invocations of forward+backward (6 distinct input sets).
  Images are independent, so N GPUs run N replicas of the per-GPU batch ("weak" scaling, no data-path collective);
  for N>1 the step also all-reduces the op's projection-weight gradients (6 x 230272 fp32) over NCCL, as DDP would.

What is timed
  value      device-resident inputs; K steps between barrier+synchronize pairs; CUDA events; max over ranks.
  roofline   per-launch CUDA-event brackets around every forward / backward kernel inside the same timed region.
  e2e        the same step through msda_host_forward_backward: pinned HOST buffers in, HOST buffers out, copies timed.
  cpu_baseline / --impl reference
             the reference's CPU path (F.grid_sample composition, restated in oracle/msda_ref_torch.py) on the host
             cores, on a bounded sample (two 800x1333 images per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (images per GPU, level shapes, Lq or None for Lq=S, M, D, L, P, dtype, layers per step)
    "detr_encoder_800x1333": dict(N=16, shapes=[(100, 167), (50, 84), (25, 42), (13, 21)], Lq=None, M=8, D=32, P=4,
                                  dtype="f32", layers=6),
    "grit_encoder_384x640": dict(N=32, shapes=[(48, 80), (24, 40), (12, 20), (6, 10)], Lq=None, M=8, D=32, P=4,
                                 dtype="f32", layers=6),
    "grit_decoder_384x640_bf16": dict(N=64, shapes=[(48, 80), (24, 40), (12, 20), (6, 10)], Lq=150, M=8, D=64, P=4,
                                      dtype="bf16", layers=6),
    "detr_encoder_800x1333_bf16": dict(N=32, shapes=[(100, 167), (50, 84), (25, 42), (13, 21)], Lq=None, M=8, D=32, P=4,
                                       dtype="bf16", layers=6),
    "detr_encoder_800x1333_d64": dict(N=8, shapes=[(100, 167), (50, 84), (25, 42), (13, 21)], Lq=None, M=8, D=64, P=4,
                                      dtype="f32", layers=6),
    "grit_decoder_800x1333_bf16": dict(N=32, shapes=[(100, 167), (50, 84), (25, 42), (13, 21)], Lq=150, M=8, D=64, P=4,
                                       dtype="bf16", layers=6),
    # GRIT's real operating point: no AMP, C=512 / D=64, 150 queries (det_module.py:285,335-336; train_config.yaml:34-41)
    "grit_decoder_384x640_f32": dict(N=64, shapes=[(48, 80), (24, 40), (12, 20), (6, 10)], Lq=150, M=8, D=64, P=4,
                                     dtype="f32", layers=6),
    "grit_decoder_800x1333_f32": dict(N=32, shapes=[(100, 167), (50, 84), (25, 42), (13, 21)], Lq=150, M=8, D=64, P=4,
                                      dtype="f32", layers=6),
    "tiny": dict(N=2, shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], Lq=None, M=8, D=32, P=4, dtype="f32", layers=2),
}
OP_PARAMS_PER_LAYER = {256: 230272, 512: 722304}  # the four Linears of one MSDeformAttn
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def algorithmic_bytes(N, S, Lq, M, D, L, P, ev, el=4):
    """Compulsory traffic of one launch (SURVEY.md section 8d / BASELINE.md section 2)."""
    V = N * S * M * D * ev
    T = N * Lq * M * L * P * 4 * D * ev
    v_eff = min(V, T)
    Lc = N * Lq * M * L * P * 2 * el
    A = N * Lq * M * L * P * el
    O = N * Lq * M * D * ev
    fwd = v_eff + Lc + A + O
    bwd = (v_eff + Lc + A + O) + (V + Lc + A)
    return fwd, bwd


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1])), power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local_rank):
    """Best effort: pin this rank's CPU threads (and therefore its first-touch pinned host buffers) to the NUMA node
    its GPU hangs off, so the e2e leg's H2D/D2H copies do not cross the socket interconnect."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out  # 00000000:1b:00.0 -> 0000:1b:00.0
        path = f"/sys/bus/pci/devices/{bus}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} CPUs local to GPU {local_rank}"
    except Exception as exc:
        return f"no NUMA binding ({type(exc).__name__})"
    return "no NUMA binding"


def make_layer_inputs(torch, cfg, device, seed, loc_dist):
    """One layer's synthetic inputs, created on `device` (SURVEY.md section 8d distributions)."""
    N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
    shapes = cfg["shapes"]
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    Lq = cfg["Lq"] or S
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f64": torch.float64}[cfg["dtype"]]
    gen = torch.Generator(device=device).manual_seed(seed)
    value = torch.randn(N, S, M, D, device=device, generator=gen).to(dt)
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, device=device, generator=gen), -1).view(N, Lq, M, L, P)
    if loc_dist == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, device=device, generator=gen) * 1.1 - 0.05
    elif loc_dist == "concentrated":  # diagnostic: every sample inside a ~2% x 2% window -> taps hit L1
        loc = 0.5 + torch.rand(N, Lq, M, L, P, 2, device=device, generator=gen) * 0.02
    else:  # detector-like: own pixel centre (encoder) or U[0,1) (decoder) + init ring offsets + N(0, 2 px)
        import math
        if Lq == S:
            refs = []
            for h, w in shapes:
                ys, xs = torch.meshgrid(torch.arange(h, device=device) + 0.5, torch.arange(w, device=device) + 0.5,
                                        indexing="ij")
                refs.append(torch.stack([xs.reshape(-1) / w, ys.reshape(-1) / h], -1))
            ref = torch.cat(refs, 0)[None].expand(N, -1, -1)
        else:
            ref = torch.rand(N, Lq, 2, device=device, generator=gen)
        ang = torch.arange(M, device=device, dtype=torch.float32) * (2.0 * math.pi / M)
        ring = torch.stack([ang.cos(), ang.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        offs = ring.view(1, 1, M, 1, 1, 2) * torch.arange(1, P + 1, device=device).view(1, 1, 1, 1, P, 1)
        offs = offs + 2.0 * torch.randn(N, Lq, M, L, P, 2, device=device, generator=gen)
        norm = torch.tensor([[w, h] for h, w in shapes], device=device, dtype=torch.float32).view(1, 1, 1, L, 1, 2)
        loc = ref.view(N, Lq, 1, 1, 1, 2) + offs / norm
    gout = torch.randn(N, Lq, M * D, device=device, generator=gen).to(dt)
    return dict(value=value.contiguous(), loc=loc.contiguous(), attn=attn.contiguous(), gout=gout.contiguous())


def load_reference_core():
    """The reference's own ``ms_deform_attn_core_pytorch`` (models/ops/functions/ms_deform_attn_func.py:41-61), loaded
    UNMODIFIED from the copy __graft_entry__.build() leaves in git-ignored baseline/_ref/ops_test.  The file imports the
    compiled extension at module top (:18); the CPU function never calls it, so an empty stand-in module satisfies the
    import.  Returns None when the copy is absent (then the restated port in oracle/msda_ref_torch.py is timed)."""
    path = os.path.join(ROOT, "baseline", "_ref", "ops_test", "functions", "ms_deform_attn_func.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    import types
    name = "MultiScaleDeformableAttention"
    prev = sys.modules.get(name)
    sys.modules[name] = types.ModuleType(name)
    try:
        spec = importlib.util.spec_from_file_location("_grit_reference_ms_deform_attn_func", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.ms_deform_attn_core_pytorch
    except Exception:
        return None
    finally:
        if prev is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = prev


def cpu_reference_pass(torch, cfg, images, threads, steps, warmup, loc_dist):
    """The reference's CPU path (grid_sample composition) fwd + autograd bwd on `images` images per step.
    Returns (queries/s, s/step, sample description, kind): kind "reference" = the reference's own function,
    "port" = the restatement in oracle/msda_ref_torch.py."""
    from oracle import msda_ref_torch
    ref_core = load_reference_core()
    torch.set_num_threads(threads)
    small = dict(cfg, N=images)
    data = make_layer_inputs(torch, dict(small, dtype="f32"), "cpu", 0, loc_dist)
    shapes = cfg["shapes"]
    S = sum(h * w for h, w in shapes)
    Lq = cfg["Lq"] or S
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if ref_core is not None:
            v = data["value"].detach().requires_grad_(True)
            lo = data["loc"].detach().requires_grad_(True)
            at = data["attn"].detach().requires_grad_(True)
            ref_core(v, shapes, lo, at).backward(data["gout"])
        else:
            msda_ref_torch.forward_backward(data["value"], shapes, data["loc"], data["attn"], data["gout"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    what = "the reference's own ms_deform_attn_core_pytorch" if ref_core is not None else "oracle/msda_ref_torch.py port"
    return images * Lq * len(times) / total, total / len(times), f"{images} image(s) of the workload shape per step, " \
        f"1 layer fwd+autograd bwd ({what}), fp32, {len(times)} timed steps after {warmup} warm-up, " \
        f"torch {torch.__version__} CPU", ("reference" if ref_core is not None else "port")


def run_reference_arm(args, cfg, rank):
    import torch
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    S = sum(h * w for h, w in cfg["shapes"])
    Lq = cfg["Lq"] or S
    qps, sec_per_step, sample, kind = cpu_reference_pass(torch, cfg, 2, threads, args.steps, args.warmup, args.loc_dist)
    line = {
        "impl": "reference", "metric": "msda_fwd_bwd_queries_per_sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, args.gpus),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, world):
    S = sum(h * w for h, w in cfg["shapes"])
    return {"workload": args.workload, "images_per_gpu": cfg["N"], "global_batch": cfg["N"] * world, "S": S,
            "Lq": cfg["Lq"] or S, "M": cfg["M"], "D": cfg["D"], "L": len(cfg["shapes"]), "P": cfg["P"],
            "level_shapes": cfg["shapes"], "layers_per_step": cfg["layers"], "pass": "forward+backward",
            "loc_dist": {"uniform": "uniform U[-0.05,1.05) (worst-case locality)",
                         "detector": "detector-like (own pixel + ring offsets + N(0,2px))",
                         "concentrated": "diagnostic: all samples in a 2% window (L1-resident taps)"}[args.loc_dist],
            "l2_policy": "inputs larger than L2: every layer reads its own input set (>= 1 GB at the default workload), "
                         "126 MB L2 is cycled between launches",
            "parallelism": f"batch-sharded replicas x{world}",
            **({"tuning": args.tuning} if args.tuning else {})}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="detr_encoder_800x1333")
    ap.add_argument("--loc-dist", choices=["uniform", "detector", "concentrated"], default="uniform")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--tuning", default="", help="A/B only: msda_set_tuning knobs for the whole run, key=value,key=value "
                                                 "(recorded in config.tuning; the line of record is run without it)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--e2e-chunk", type=int, default=None, help="images per pipeline chunk of the host-buffer path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:  # convenience: self-launch one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                                   os.environ.get("MASTER_PORT", "29541"), os.path.abspath(__file__)] + sys.argv[1:])

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    from grit_b200 import _lib

    # multi-rank: bind each rank (and so its first-touch pinned buffers) to its GPU's NUMA node unless told not to
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 and os.environ.get("MSDA_BENCH_NUMA_BIND", "1") == "1" \
        else None
    lib = _lib.load()  # raises if the CUDA library is missing: no fallback
    for kv in filter(None, args.tuning.split(",")):
        _lib.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
    L = len(cfg["shapes"])
    S = sum(h * w for h, w in cfg["shapes"])
    Lq = cfg["Lq"] or S
    layers = cfg["layers"]
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}[cfg["dtype"]]
    ev = 4 if cfg["dtype"] == "f32" else 2
    shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=device)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    sets = [make_layer_inputs(torch, cfg, device, 1000 * rank + i, args.loc_dist) for i in range(layers)]
    out = torch.empty(N, Lq, M * D, device=device, dtype=dt)
    gv = torch.empty(N, S, M, D, device=device, dtype=dt)
    gl = torch.empty(N, Lq, M, L, P, 2, device=device)
    ga = torch.empty(N, Lq, M, L, P, device=device)
    import ctypes
    dims = _lib.MsdaDims(N, S, M, D, L, Lq, P)
    code = _lib._DTYPE_CODE[dt]
    ws_bytes = lib.msda_backward_workspace_bytes(ctypes.byref(dims), code, 0)
    ws = torch.empty(max(ws_bytes // 4, 4), dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream().cuda_stream
    P_ = _lib._ptr
    d_model = M * D
    n_bucket = layers * OP_PARAMS_PER_LAYER.get(d_model, 4 * d_model * d_model)
    grad_bucket = torch.zeros(n_bucket, device=device) if world > 1 else None
    # what each rank contributes to the all-reduce: a rank-dependent, non-zero pattern whose sum over ranks is known
    grad_src = ((torch.arange(n_bucket, device=device) % 251).float() * 1e-3 + 1.0) * (rank + 1) if world > 1 else None

    def multi_gpu_selfcheck():
        """Untimed, world > 1: (1) every rank runs ONE layer on identical inputs (shared seed) with the deterministic
        backward; bit-level checksums of all four results are all-gathered and must be equal on every rank -- the
        replicas compute the same function.  (2) the bucket all-reduce of a non-zero rank-dependent pattern must equal
        the closed-form sum."""
        chk_cfg = dict(cfg, N=min(N, 2))
        x = make_layer_inputs(torch, chk_cfg, device, 424242, args.loc_dist)  # same seed on every rank
        o = _lib.forward(x["value"], shapes, lsi, x["loc"], x["attn"])
        g3 = _lib.backward(x["value"], shapes, lsi, x["loc"], x["attn"], x["gout"], _lib.FLAG_DETERMINISTIC)

        def digest(t):
            bits = t.contiguous().view(torch.int16 if t.element_size() == 2 else torch.int32)
            return bits.to(torch.int64).sum()
        mine = torch.stack([digest(t) for t in (o,) + tuple(g3)])
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        same = all(bool(torch.equal(g, gathered[0])) for g in gathered)
        grad_bucket.copy_(grad_src)
        dist.all_reduce(grad_bucket)
        expect = grad_src / (rank + 1) * (world * (world + 1) / 2)
        sum_ok = bool(torch.allclose(grad_bucket, expect, rtol=1e-6, atol=0))
        if not same or not sum_ok:
            raise RuntimeError(f"multi-GPU self-check failed on rank {rank}: ranks_bit_identical={same}, "
                               f"allreduce_sum_ok={sum_ok}")
        return {"ranks_bit_identical": same, "allreduce_sum_ok": sum_ok,
                "what": "one layer on identical inputs per rank, deterministic backward, int checksums of out/grad_value/"
                        "grad_loc/grad_attn all-gathered and compared; bucket all-reduce of a rank-dependent pattern "
                        "checked against its closed form"}

    def fwd(s):
        rc = lib.msda_forward(P_(s["value"]), P_(shapes), P_(lsi), P_(s["loc"]), P_(s["attn"]), P_(out),
                              ctypes.byref(dims), code, 0, ctypes.c_void_p(stream))
        if rc:
            raise RuntimeError(lib.msda_last_error().decode())

    # strategy 3 (owned) writes every grad_value line once and needs no zero-fill; the others accumulate into a zeroed
    # grad_value (fp32: zero_() inside the step, outside the kernel bracket, as in round 1; bf16: the library's workspace)
    owned = lib.msda_backward_strategy(ctypes.byref(dims), code, 0) == 3

    def bwd(s):
        flags = _lib.FLAG_ZERO_GRAD_VALUE if (dt == torch.bfloat16 or owned) else 0
        rc = lib.msda_backward(P_(s["value"]), P_(shapes), P_(lsi), P_(s["loc"]), P_(s["attn"]), P_(s["gout"]),
                               P_(gv), P_(gl), P_(ga), ctypes.byref(dims), code, flags, P_(ws), ws_bytes,
                               ctypes.c_void_p(stream))
        if rc:
            raise RuntimeError(lib.msda_last_error().decode())

    brackets = []  # (kind, start_event, end_event) for every kernel launch in the timed region

    def step(record):
        handles = []
        for li, s in enumerate(sets):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            fwd(s)
            if record:
                e1.record()
                brackets.append(("fwd", e0, e1))
            if dt != torch.bfloat16 and not owned:
                gv.zero_()  # grad_value is accumulated into: the zero-fill is compulsory work of the step
            if record:
                e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e2.record()
            bwd(s)
            if record:
                e3.record()
                brackets.append(("bwd", e2, e3))
            if grad_bucket is not None:  # pack this layer's (synthetic, non-zero) projection gradients, reduce them async
                n = grad_bucket.numel() // layers
                grad_bucket[li * n:(li + 1) * n].copy_(grad_src[li * n:(li + 1) * n])
                handles.append(dist.all_reduce(grad_bucket[li * n:(li + 1) * n], async_op=True))
        for h in handles:
            h.wait()

    for _ in range(args.warmup):
        step(False)
    barrier()
    mgpu_check = multi_gpu_selfcheck() if world > 1 else None
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.msda_launch_count(1)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    barrier()
    launches = int(lib.msda_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    if grad_bucket is not None:  # the timed region's last all-reduce summed what it should have
        expect = grad_src / (rank + 1) * (world * (world + 1) / 2)
        if not torch.allclose(grad_bucket, expect, rtol=1e-6, atol=0):
            raise RuntimeError("all-reduce inside the timed region produced a wrong sum")
    elapsed_ms = t0.elapsed_time(t1)
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    fwd_ms = [a.elapsed_time(b) for k, a, b in brackets if k == "fwd"]
    bwd_ms = [a.elapsed_time(b) for k, a, b in brackets if k == "bwd"]
    kernel_names = (None, None)
    fwd(sets[0]); kf = _lib.last_kernel(); bwd(sets[0]); kb = _lib.last_kernel(); torch.cuda.synchronize()
    kernel_names = (kf, kb)

    queries_per_step = layers * N * Lq * world
    value_qps = queries_per_step * args.steps / (elapsed_ms * 1e-3)
    fwd_b, bwd_b = algorithmic_bytes(N, S, Lq, M, D, L, P, ev)
    peak, peak_src = hbm_peak()
    avg_f, avg_b = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get(kb)
        except Exception:
            traffic = None
    med = lambda xs: sorted(xs)[len(xs) // 2]
    roof_b = {"kernel": kb, "bound": "hbm", "achieved": bwd_b / (avg_b * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
              "frac": bwd_b / (avg_b * 1e-3) / 1e9 / peak, "traffic": traffic, "algorithmic_bytes": bwd_b,
              "avg_launch_ms": avg_b, "median_launch_ms": med(bwd_ms), "min_launch_ms": min(bwd_ms),
              "launches_timed": len(bwd_ms), "peak_source": peak_src}
    ftraffic = None
    if os.path.exists(tpath):
        try:
            ftraffic = json.load(open(tpath)).get(args.workload, {}).get(kf)
        except Exception:
            ftraffic = None
    roof_f = {"kernel": kf, "bound": "hbm", "achieved": fwd_b / (avg_f * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
              "frac": fwd_b / (avg_f * 1e-3) / 1e9 / peak, "traffic": ftraffic, "algorithmic_bytes": fwd_b,
              "avg_launch_ms": avg_f, "median_launch_ms": med(fwd_ms), "min_launch_ms": min(fwd_ms),
              "launches_timed": len(fwd_ms)}
    step_gbs = (fwd_b + bwd_b) * layers * args.steps / (elapsed_ms * 1e-3) / 1e9  # per GPU

    # ---- end to end: pinned host buffers through the C ABI's host entry point --------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or max(2, min(args.steps, 5))
        host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in sets[0].items()}
        for k in host:
            host[k].copy_(sets[0][k])
        h_out = torch.empty(out.shape, dtype=dt).pin_memory()
        h_gv = torch.empty(gv.shape, dtype=dt).pin_memory()
        h_gl = torch.empty(gl.shape, dtype=torch.float32).pin_memory()
        h_ga = torch.empty(ga.shape, dtype=torch.float32).pin_memory()
        h_shapes, h_lsi = shapes.cpu(), lsi.cpu()
        sess = _lib.HostSession(dims, dt, device=local_rank, images_per_chunk=args.e2e_chunk or max(1, N // 16))

        def e2e_step():  # the six layers are submitted back to back so the copy/kernel pipeline never drains between them
            for _ in range(layers):
                sess.submit(host["value"], h_shapes, h_lsi, host["loc"], host["attn"], host["gout"], h_out,
                            h_gv, h_gl, h_ga)
            sess.wait()
        e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - w0
        if world > 1:
            tmax = torch.tensor([e2e_s], device=device)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_s = float(tmax.item())
        h2d = sum(v.numel() * v.element_size() for v in host.values()) * layers
        d2h = sum(t.numel() * t.element_size() for t in (h_out, h_gv, h_gl, h_ga)) * layers
        e2e = {"value": queries_per_step * e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
               "path": "msda_host_submit x layers + msda_host_wait (pinned host buffers, chunked H2D/compute/D2H "
                       "pipeline kept full across the layers); device->host read = all four result tensors",
               "numa": numa_note}
        sess.close()
        # host ceiling: the same bytes with plain copies only (no kernels), H2D and D2H on two streams, all ranks at once
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dev_in = {k: torch.empty_like(v) for k, v in sets[0].items()}

        def copy_step():
            for _ in range(layers):
                with torch.cuda.stream(s_in):
                    for k in host:
                        dev_in[k].copy_(host[k], non_blocking=True)
                with torch.cuda.stream(s_out):
                    h_out.copy_(out, non_blocking=True), h_gv.copy_(gv, non_blocking=True)
                    h_gl.copy_(gl, non_blocking=True), h_ga.copy_(ga, non_blocking=True)
            s_in.synchronize(), s_out.synchronize()
        copy_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(2):
            copy_step()
        barrier()
        copy_s = (time.perf_counter() - w0) / 2
        if world > 1:
            tmax = torch.tensor([copy_s], device=device)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            copy_s = float(tmax.item())
        e2e["copy_only_ms_per_step"] = copy_s * 1e3
        e2e["copy_only_gbs_per_gpu"] = (h2d + d2h) / copy_s / 1e9
        e2e["frac_of_copy_ceiling"] = copy_s / (e2e_s / e2e_steps)
        e2e["ceiling_note"] = "copy_only = the step's H2D + D2H bytes with cudaMemcpyAsync alone (no kernels), both " \
                              "directions concurrently, all ranks at once: the host/PCIe ceiling of this path"
        del dev_in
        # cheap sanity: the host path and the device path computed the same thing for the last layer set
        fwd(sets[0]); torch.cuda.synchronize()
        if not torch.equal(out.cpu(), h_out):
            raise RuntimeError("e2e host path and device path disagree")

    # ---- on-chip ceilings: what actually bounds a gather/scatter whose tap traffic is ~18x its HBM traffic -------------
    if rank == 0 and not args.no_extras and dt == torch.float32 and D == 32:
        region = torch.zeros(S * M * D, dtype=torch.float32, device=device)  # one image of value: L2-resident
        taps = N * Lq * M * L * P * 4  # 128-byte tap lines per launch (forward gathers them, backward also reduces them)
        for roof, which, avg_ms in ((roof_f, "gather", avg_f), (roof_b, "red", avg_b)):
            ceiling = _lib.probe_ceiling(which, region)
            achieved = taps / (avg_ms * 1e-3) / 1e9
            roof["on_chip"] = {
                "resource": {"gather": "L2->L1 gather of random 128-byte lines (LDG.E.128 stream)",
                             "red": "L2 reduction of random 128-byte lines (REDG.E.ADD.F32x4 stream)"
                                    + ("; bwd_planes sends the smallest levels' taps to shared-memory integer atomics "
                                       "instead, so its tap rate is not capped by this ceiling"
                                       if (kernel_names[1] or "").startswith("bwd_planes") else "")}[which],
                "ceiling_glines_per_s": ceiling, "achieved_glines_per_s": achieved, "frac": achieved / ceiling,
                "how": "msda_probe_ceiling microbenchmark, same access pattern, no arithmetic, measured in this run"}
        del region

    # ---- extras (rank 0, N=1): the other location distribution and the reference's own CUDA kernels on this GPU -------
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}

        def graph_pair(inputs, fwd_fn, bwd_fn, calls=10, replays=10):
            """Per-call time of fwd_fn / bwd_fn when `calls` invocations are captured in one CUDA graph and replayed."""
            res = []
            for fn in (fwd_fn, bwd_fn):
                fn(inputs)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for _ in range(calls):
                        fn(inputs)
                graph.replay()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(replays):
                    graph.replay()
                b.record()
                torch.cuda.synchronize()
                res.append(a.elapsed_time(b) / (calls * replays))
                del graph
            return res

        def time_pair(inputs, fwd_fn, bwd_fn, iters=5):
            res = []
            for fn in (fwd_fn, bwd_fn):
                for _ in range(2):
                    fn(inputs)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    fn(inputs)
                b.record()
                torch.cuda.synchronize()
                res.append(a.elapsed_time(b) / iters)
            return res

        other = "detector" if args.loc_dist == "uniform" else "uniform"
        alt = make_layer_inputs(torch, cfg, device, 77, other)

        def bwd_zero(s):
            if dt != torch.bfloat16 and not owned:
                gv.zero_()
            bwd(s)
        f_ms, b_ms = time_pair(alt, fwd, bwd_zero)
        extras[f"loc_dist_{other}"] = {"fwd_ms": f_ms, "bwd_ms": b_ms, "queries_per_s": N * Lq / ((f_ms + b_ms) * 1e-3)}
        del alt
        # the other BASELINE.json configs, one quick forward/backward timing each (public API, device tensors)
        for name in ("grit_encoder_384x640", "grit_decoder_384x640_bf16", "grit_decoder_800x1333_bf16"):
            if name == args.workload:
                continue
            c2 = WORKLOADS[name]
            x = make_layer_inputs(torch, c2, device, 5, "uniform")
            sh = torch.tensor(c2["shapes"], dtype=torch.int64, device=device)
            ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
            f_ms, b_ms = time_pair(x, lambda s_: _lib.forward(s_["value"], sh, ls, s_["loc"], s_["attn"]),
                                   lambda s_: _lib.backward(s_["value"], sh, ls, s_["loc"], s_["attn"], s_["gout"]))
            S2 = sum(h * w for h, w in c2["shapes"])
            Lq2 = c2["Lq"] or S2
            fb, bb = algorithmic_bytes(c2["N"], S2, Lq2, c2["M"], c2["D"], len(c2["shapes"]), c2["P"],
                                       4 if c2["dtype"] == "f32" else 2)
            extras[name] = {"dtype": c2["dtype"], "N": c2["N"], "Lq": Lq2, "S": S2, "D": c2["D"],
                            "fwd_ms": f_ms, "bwd_ms_incl_alloc_zero_fold": b_ms,
                            "fwd_queries_per_s": c2["N"] * Lq2 / (f_ms * 1e-3),
                            "fwd_bwd_queries_per_s": c2["N"] * Lq2 / ((f_ms + b_ms) * 1e-3),
                            "fwd_hbm_frac": fb / (f_ms * 1e-3) / 1e9 / peak, "bwd_hbm_frac": bb / (b_ms * 1e-3) / 1e9 / peak}
            del x
        # GRIT's real operating point (no AMP: fp32, 150 queries, C=512 / D=64; det_module.py:285,335-336) at batch 4 / 16 /
        # 64, with the HBM-roofline fraction of each direction and the reference's own CUDA kernels beside them
        ref_so = os.path.join(ROOT, "baseline", "_ref", "MultiScaleDeformableAttentionRef.so")
        refmod = None
        if os.path.exists(ref_so):
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttentionRef", ref_so)
                refmod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(refmod)
            except Exception:
                refmod = None
        for name in ("grit_decoder_384x640_f32", "grit_decoder_800x1333_f32"):
            for nb in (4, 16, 64):
                c2 = dict(WORKLOADS[name], N=nb)
                x = make_layer_inputs(torch, c2, device, 5, "uniform")
                sh = torch.tensor(c2["shapes"], dtype=torch.int64, device=device)
                ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
                f_ms, b_ms = time_pair(x, lambda s_: _lib.forward(s_["value"], sh, ls, s_["loc"], s_["attn"]),
                                       lambda s_: _lib.backward(s_["value"], sh, ls, s_["loc"], s_["attn"], s_["gout"]),
                                       iters=20)
                S2 = sum(h * w for h, w in c2["shapes"])
                fb, bb = algorithmic_bytes(nb, S2, c2["Lq"], c2["M"], c2["D"], len(c2["shapes"]), c2["P"], 4)
                entry = {"fwd_us": f_ms * 1e3, "bwd_us_incl_alloc_zero": b_ms * 1e3,
                         "fwd_hbm_frac": fb / (f_ms * 1e-3) / 1e9 / peak, "bwd_hbm_frac": bb / (b_ms * 1e-3) / 1e9 / peak,
                         "fwd_bwd_queries_per_s": nb * c2["Lq"] / ((f_ms + b_ms) * 1e-3)}
                # the same calls replayed from a CUDA graph: at batch 4 a call is 5-15 us of kernel behind 20-30 us of Python
                # and launch overhead, so only the graphed figures compare KERNELS (GRIT's decoder would be graphed too)
                gf = gb = None
                try:
                    gf, gb = graph_pair(x, lambda s_: _lib.forward(s_["value"], sh, ls, s_["loc"], s_["attn"]),
                                        lambda s_: _lib.backward(s_["value"], sh, ls, s_["loc"], s_["attn"], s_["gout"]))
                    entry.update(graphed_fwd_us=gf * 1e3, graphed_bwd_us_incl_alloc_zero=gb * 1e3)
                except Exception as exc:
                    entry["graphed"] = repr(exc)[:120]
                if refmod is not None:
                    try:
                        ref_f = lambda s_: refmod.ms_deform_attn_forward(s_["value"], sh, ls, s_["loc"], s_["attn"], 64)
                        ref_b = lambda s_: refmod.ms_deform_attn_backward(s_["value"], sh, ls, s_["loc"], s_["attn"],
                                                                          s_["gout"], 64)
                        rf, rb = time_pair(x, ref_f, ref_b, iters=20)
                        entry.update(ref_cuda_fwd_us=rf * 1e3, ref_cuda_bwd_us=rb * 1e3,
                                     speedup_vs_ref_cuda=(rf + rb) / (f_ms + b_ms))
                        if gf is not None:
                            rgf, rgb = graph_pair(x, ref_f, ref_b)
                            entry.update(ref_cuda_graphed_fwd_us=rgf * 1e3, ref_cuda_graphed_bwd_us=rgb * 1e3,
                                         speedup_vs_ref_cuda_graphed=(rgf + rgb) / (gf + gb))
                    except Exception as exc:
                        entry["ref_cuda"] = repr(exc)[:120]
                extras[f"{name}_N{nb}"] = entry
                del x
        # the module around the op (4 Linears + pre-op arithmetic + op), reference-shaped path vs fused kernels (8f-1)
        if dt == torch.float32:
            try:
                from grit_b200 import MSDeformAttn
                nm = min(N, 4)
                torch.manual_seed(0)
                mod = MSDeformAttn(M * D, L, M, P).to(device)
                mod.validate_shapes = False
                with torch.no_grad():
                    mod.sampling_offsets.weight.normal_(0, 0.01)
                    mod.attention_weights.weight.normal_(0, 0.1)
                mq = torch.randn(nm, Lq, M * D, device=device, requires_grad=True)
                ms_ = torch.randn(nm, S, M * D, device=device, requires_grad=True)
                mr = torch.rand(nm, Lq, L, 2, device=device)
                mg = torch.randn(nm, Lq, M * D, device=device)
                mm = torch.zeros(nm, S, dtype=torch.bool, device=device)
                mm[:, ::10] = True
                res = {}
                for fused in (False, True):
                    mod.fused = fused

                    def mstep(_):
                        mq.grad = ms_.grad = None
                        mod.zero_grad(set_to_none=True)
                        mod(mq, mr, ms_, shapes, lsi, mm).backward(mg)
                    for _ in range(2):
                        mstep(None)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(3):
                        mstep(None)
                    b.record()
                    torch.cuda.synchronize()
                    res["fused" if fused else "reference_shaped"] = a.elapsed_time(b) / 3
                extras["module_fwd_bwd_ms"] = dict(res, images=nm, note="MSDeformAttn module incl. its four fp32 Linears "
                                                   "(cuBLAS, torch default precision), padding mask on 10% of pixels")
                del mod, mq, ms_, mr, mg, mm
            except Exception as exc:
                extras["module_fwd_bwd_ms"] = {"unavailable": repr(exc)[:200]}
        if refmod is not None and dt == torch.float32:
            try:
                s0 = sets[0]
                f_ms, b_ms = time_pair(
                    s0, lambda s: refmod.ms_deform_attn_forward(s["value"], shapes, lsi, s["loc"], s["attn"], 64),
                    lambda s: refmod.ms_deform_attn_backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"],
                                                             64))
                mine = (avg_f + avg_b)
                extras["reference_cuda_kernels_on_this_gpu"] = {
                    "what": "the reference's own CUDA kernels (models/ops/src/cuda) recompiled for sm_100a "
                            "(baseline/build_ref_cuda.py), same inputs, incl. their output allocation + zero-fill",
                    "fwd_ms": f_ms, "bwd_ms": b_ms, "queries_per_s": N * Lq / ((f_ms + b_ms) * 1e-3),
                    "speedup_of_this_repo_kernels": (f_ms + b_ms) / mine}
            except Exception as exc:  # the comparison is optional; never fail the bench because of it
                extras["reference_cuda_kernels_on_this_gpu"] = {"unavailable": repr(exc)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, _, sample, kind = cpu_reference_pass(torch, cfg, 2, threads, 16, 1, args.loc_dist)  # ~10 s of CPU work
        cpu = {"value": v, "unit": "queries/s", "cores": threads, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": "msda_fwd_bwd_queries_per_sec", "value": value_qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"],
            "data": "synthetic", "config": workload_config(args, cfg, world),
            "hbm_gbs_per_gpu": step_gbs, "hbm_frac_step": step_gbs / peak,
            "roofline": roof_b, "roofline_fwd": roof_f, "cpu_baseline": cpu, "e2e": e2e, "extras": extras,
            "gpu_launches": launches, "kernels": {"forward": kernel_names[0], "backward": kernel_names[1]},
            "multi_gpu_check": mgpu_check, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
