"""GPU experiment: does an L2 access-policy persistence window help the op (north_star: "the finest level gets an L2
access-policy persistence window")?  Measures forward and backward with and without a persisting window over `value`.

    python scripts/l2_window_experiment.py [--out gpurun_out/r2_l2_window.json]

Setup per workload: cudaLimitPersistingL2CacheSize = the device maximum, then a stream access-policy window
(hitProp = persisting, missProp = streaming) over the first min(window max, 100 MB, tensor) bytes of `value` -- image 0..k
in full, which contains those images' finest levels -- on the stream the kernels run on.  The same inputs are re-run back
to back, so whatever reuse a window can create is present (in training every layer has its own `value`, so this is an
upper bound).  Result (B200, recorded in DESIGN.md section 5): no measurable effect -- the encoder shapes already hit in L2
(one image's maps are 23 MB against 126 MB of L2, rows are processed image-major) and the decoder shapes read each
`value` line at most a few times within one launch.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from grit_b200 import _lib  # noqa: E402


def set_window(stream_handle, ptr, nbytes, hit_ratio):
    from cuda import cudart
    attr = getattr(cudart, "cudaStreamAttrValue", None) or cudart.cudaLaunchAttributeValue
    attr = attr()
    attr.accessPolicyWindow.base_ptr = ptr
    attr.accessPolicyWindow.num_bytes = nbytes
    attr.accessPolicyWindow.hitRatio = hit_ratio
    attr.accessPolicyWindow.hitProp = cudart.cudaAccessProperty.cudaAccessPropertyPersisting
    attr.accessPolicyWindow.missProp = cudart.cudaAccessProperty.cudaAccessPropertyStreaming
    ids = cudart.cudaStreamAttrID
    attr_id = getattr(ids, "cudaStreamAttributeAccessPolicyWindow", None) or ids.cudaLaunchAttributeAccessPolicyWindow
    err, = cudart.cudaStreamSetAttribute(stream_handle, attr_id, attr)
    return int(err)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from cuda import cudart
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.load()
    err, prop = cudart.cudaGetDeviceProperties(0)
    max_persist, max_window = int(prop.persistingL2CacheMaxSize), int(prop.accessPolicyMaxWindowSize)
    err, = cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitPersistingL2CacheSize, max_persist)
    results = {"persistingL2CacheMaxSize": max_persist, "accessPolicyMaxWindowSize": max_window, "l2_bytes": int(prop.l2CacheSize),
               "set_limit_err": int(err)}
    stream = torch.cuda.Stream()
    for name in ("detr_encoder_800x1333", "grit_decoder_800x1333_f32", "grit_decoder_800x1333_bf16",
                 "grit_decoder_384x640_f32"):
        cfg = dict(bench.WORKLOADS[name])
        x = bench.make_layer_inputs(torch, cfg, dev, 3, "uniform")
        shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        res = {}
        with torch.cuda.stream(stream):
            def fwd():
                return _lib.forward(x["value"], shapes, lsi, x["loc"], x["attn"])

            def bwd():
                return _lib.backward(x["value"], shapes, lsi, x["loc"], x["attn"], x["gout"])

            def timeit(fn):
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(args.iters):
                    fn()
                b.record(stream)
                stream.synchronize()
                return a.elapsed_time(b) / args.iters
            nbytes = min(x["value"].numel() * x["value"].element_size(), max_window, 100 << 20)
            for label, ratio in (("no_window", 0.0), ("window", 1.0), ("no_window_again", 0.0)):
                rc = set_window(stream.cuda_stream, x["value"].data_ptr(), nbytes if ratio > 0 else 0, ratio)
                res[label] = {"fwd_ms": timeit(fwd), "bwd_ms": timeit(bwd), "set_attr_err": rc, "window_bytes": nbytes if ratio > 0 else 0}
        cudart.cudaCtxResetPersistingL2Cache()
        results[name] = res
        print(name, json.dumps(res), flush=True)
        del x
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
