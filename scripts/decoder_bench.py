"""GPU: GRIT's real operating point -- the six-layer deformable decoder, 150 queries, C=512 (D=64), fp32, no AMP
(reference: models/detection/det_module.py:181-211, 313-349; configs/detection/train_config.yaml:34-41), and the
feature-extraction consumer (tools/extract_features.py: batch 64, forward only).

    python scripts/decoder_bench.py [--pyramid 384x640|800x1333] [--batches 4,16,64] [--out file.json]

For every batch size it reports, for the SAME weights and inputs:
  reference_launches  the layer as the reference launches it: per-layer value_proj, softmax / location arithmetic /
                      masked_fill as separate PyTorch kernels around the sampling op, dropout + add + LayerNorm epilogues
  fused_eager         hoisted value_proj (one GEMM + one mask fill for six layers), fused sampling kernels with in-kernel
                      valid-ratio scaling, fused residual + dropout + LayerNorm epilogues; fused_eager_no_hoist: the same with
                      per-layer value_proj (what hoist_value_proj="auto" picks above 256 MB of values; the training step and
                      the graphed decoder use "auto")
  fused_graphed       the same under ONE CUDA graph (GraphedDecoder): six-layer forward latency, images/s
and a training step (forward + backward of the six layers) for the first two.  Kernel launches per layer are counted
with torch.profiler (CUPTI) when it is available.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import grit_b200  # noqa: E402
from grit_b200 import DeformableTransformerDecoderLayer, GraphedDecoder, run_decoder  # noqa: E402

PYRAMIDS = {"384x640": [(48, 80), (24, 40), (12, 20), (6, 10)], "800x1333": [(100, 167), (50, 84), (25, 42), (13, 21)]}


def count_launches(fn):
    try:
        from torch.profiler import ProfilerActivity, profile
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        n = sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)
        return n if n > 0 else None
    except Exception:
        return None


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pyramid", default="384x640")
    ap.add_argument("--batches", default="4,16,64")
    ap.add_argument("--d-model", type=int, default=512)
    ap.add_argument("--out", default=None)
    ap.add_argument("--matmul", default="fp32", choices=["fp32", "tf32"],
                    help="precision of the layer's cuBLAS GEMMs: fp32 = torch's default, what GRIT runs (no AMP, TF32 off); "
                         "tf32 = torch.backends.cuda.matmul.allow_tf32, a user-side switch outside this library that "
                         "shows what the layer costs once the GEMMs stop dominating it")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = args.matmul == "tf32"
    shapes_l = PYRAMIDS[args.pyramid]
    C, M, L, P, Lq, n_layers = args.d_model, 8, 4, 4, 150, 6
    S = sum(h * w for h, w in shapes_l)
    dev = "cuda"
    torch.manual_seed(0)
    layers = torch.nn.ModuleList([DeformableTransformerDecoderLayer(C, 1024, 0.1, "relu", L, M, P)
                                  for _ in range(n_layers)]).to(dev)
    for layer in layers:
        layer.cross_attn.validate_shapes = False
        with torch.no_grad():
            layer.cross_attn.sampling_offsets.weight.normal_(0, 0.02)
            layer.cross_attn.attention_weights.weight.normal_(0, 0.2)
    shapes = torch.tensor(shapes_l, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    results = {"pyramid": args.pyramid, "S": S, "Lq": Lq, "d_model": C, "layers": n_layers, "gpu": torch.cuda.get_device_name(0),
               "matmul": args.matmul}

    def configure(fused):
        for layer in layers:
            layer.fused_epilogue = fused
            layer.cross_attn.fused = fused

    for N in [int(b) for b in args.batches.split(",")]:
        tgt, pos = torch.randn(N, Lq, C, device=dev), torch.randn(N, Lq, C, device=dev)
        ref = torch.rand(N, Lq, 2, device=dev)
        src = torch.randn(N, S, C, device=dev)
        vr = torch.rand(N, L, 2, device=dev) * 0.2 + 0.8
        mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
        mask[:, ::10] = True
        res = {}

        def fwd(hoist):
            with torch.no_grad():
                return run_decoder(layers, tgt, pos, ref, src, shapes, lsi, vr, mask, hoist_value_proj=hoist)

        def train_step(hoist):
            for p_ in layers.parameters():
                p_.grad = None
            t = tgt.clone().requires_grad_(True)
            out = run_decoder(layers, t, pos, ref, src, shapes, lsi, vr, mask, hoist_value_proj=hoist)
            out[-1].sum().backward()

        layers.eval()
        configure(False)
        res["reference_launches"] = {"fwd_ms": timeit(lambda: fwd(False)),
                                     "launches_per_layer_fwd": None if (n := count_launches(lambda: fwd(False))) is None else n / n_layers}
        ref_out = fwd(False)
        configure(True)
        res["fused_eager"] = {"fwd_ms": timeit(lambda: fwd(True)),
                              "launches_per_layer_fwd": None if (n := count_launches(lambda: fwd(True))) is None else n / n_layers}
        res["fused_eager"]["max_err_vs_reference_launches"] = float((fwd(True) - ref_out).abs().max() / ref_out.abs().max())
        res["fused_eager_no_hoist"] = {"fwd_ms": timeit(lambda: fwd(False))}  # fused kernels, per-layer value_proj
        graphed = GraphedDecoder(layers, tgt, pos, ref, src, shapes, lsi, vr, mask)  # hoist_value_proj="auto"
        ms = timeit(lambda: graphed(tgt, pos, ref, src, vr, mask))
        res["fused_graphed"] = {"fwd_ms": ms, "six_layer_latency_us": ms * 1e3, "images_per_s": N / (ms * 1e-3),
                                "launches": 1}
        del graphed
        layers.train()
        configure(False)
        res["reference_launches"]["train_step_ms"] = timeit(lambda: train_step(False), iters=10, warmup=3)
        configure(True)
        res["fused_eager"]["train_step_ms"] = timeit(lambda: train_step("auto"), iters=10, warmup=3)
        try:  # forward and backward of the stack replayed from CUDA graphs (torch.cuda.make_graphed_callables)
            from grit_b200 import graphed_training_decoder
            gt, gs = tgt.clone().requires_grad_(True), src.clone()
            gfn = graphed_training_decoder(layers, gt, pos, ref, gs, shapes, lsi, vr, mask)

            def graphed_step():
                for p_ in layers.parameters():
                    p_.grad = None
                t = tgt.clone().requires_grad_(True)
                gfn(t, pos, ref, src, shapes, lsi, vr, mask)[-1].sum().backward()
            res["fused_graphed_training"] = {"fwd_ms": 0.0, "train_step_ms": timeit(graphed_step, iters=10, warmup=3)}
            del gfn
        except Exception as exc:  # noqa: BLE001
            res["fused_graphed_training"] = {"fwd_ms": 0.0, "train_step_ms": 0.0, "error": repr(exc)[:200]}
        res["speedup_fwd_graphed_vs_reference_launches"] = res["reference_launches"]["fwd_ms"] / res["fused_graphed"]["fwd_ms"]
        res["speedup_train_step"] = res["reference_launches"]["train_step_ms"] / res["fused_eager"]["train_step_ms"]
        results[f"N{N}"] = res
        print(f"N={N}", json.dumps(res), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
