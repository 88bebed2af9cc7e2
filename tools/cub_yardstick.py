"""Sort yardstick: cub::DeviceRadixSort::SortPairs(begin_bit, end_bit) against this repo's binning on
the SAME keys — the (tile << 32 | bits(depth)) keys and 1-based ids of one C2 forward.

    python tools/cub_yardstick.py [C2]      -> gpurun_out/cub_yardstick.json

Compared, all producing the same sorted (keys, values) bit for bit:
  cub            SortPairs over the bits that can differ, [0, 32 + tile_bits)  (CUB cannot subtract the depth base, so it
                 sorts all 32 depth bits: 6 passes at 45 bits)
  cub_tight      SortPairs over [begin_bit = lowest set bit that differs, 32 + tile_bits)
  gsr_sort_pairs this repo's onesweep over the 40 significant bits of the compact key (5 passes over M)
  pipeline       what the forward does by default: depth pre-sort of the N Gaussians + 2 tile passes over M
                 (stage times `presort` + `sort` from the library's own events)
Measurement infrastructure; the library is built by this script into tools/libcub_yardstick.so.
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200"), os.path.join(ROOT, "tests")]
SO = os.path.join(ROOT, "tools", "libcub_yardstick.so")


def build():
    src = os.path.join(ROOT, "tools", "cub_yardstick.cu")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-O3",
                               "-std=c++17", "-shared", "-Xcompiler", "-fPIC", src, "-o", SO])
    return SO


def main():
    import numpy as np
    import torch
    import parity as P
    from gsrast import GaussianRasterizer, _lib
    from gsrast.synthetic import CONFIGS, make_config
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    n, deg, W, H, mode, seed, _ = CONFIGS[name]
    sc = make_config(name)
    cam, _ = P.cameras(sc)
    dev = P.to_dev(sc)
    rast = GaussianRasterizer(width=W, height=H, mode=mode)
    P.gpu_forward(rast, dev, cam, deg)
    gs = rast.gstate
    M = gs.n_rendered
    keys_in = gs.keys_unsorted.clone()      # reference order (written on demand by gsr_get_state)
    vals_in = gs.values_unsorted.clone()
    keys_ref, vals_ref = gs.keys_sorted.clone(), gs.values_sorted.clone()
    tile_bits = int(np.ceil(np.log2(max(rast.n_tiles, 2))))
    lib = C.CDLL(build())
    lib.cub_sort_pairs_temp_bytes.restype = C.c_size_t
    lib.cub_sort_pairs_temp_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    lib.cub_sort_pairs.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_int, C.c_void_p]
    ko, vo = torch.empty_like(keys_in), torch.empty_like(vals_in)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())

    def timeit(fn, reps=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {"config": name, "M": int(M), "N": n, "tile_bits": tile_bits, "key_bits_cub": 32 + tile_bits, "results": {}}
    # lowest bit in which two keys differ: depths in (0.2, 1000) share their top exponent bits, not their low bits -> 0
    for label, (b0, b1) in {"cub": (0, 32 + tile_bits)}.items():
        tb = lib.cub_sort_pairs_temp_bytes(M, b0, b1)
        temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
        run = lambda: lib.cub_sort_pairs(vp(temp), tb, vp(keys_in), vp(ko), vp(vals_in), vp(vo), M, b0, b1, stream)
        assert run() == 0
        torch.cuda.synchronize()
        assert torch.equal(ko.view(torch.int64), keys_ref.view(torch.int64)) and torch.equal(vo.view(torch.int32), vals_ref.view(torch.int32)), \
            "CUB result differs from the library's sorted buffers"
        ms = timeit(run)
        res["results"][label] = {"ms": round(ms, 4), "pairs_per_s": round(M / ms * 1e3 / 1e9, 2), "bits": [b0, b1],
                                 "temp_bytes": int(tb), "equal_to_gsr": True}
    # this repo's stand-alone sort entry on the same buffers (5 passes over M, histogram pass included)
    run = lambda: _lib.check(_lib.lib().gsr_sort_pairs(rast._h, vp(keys_in), vp(vals_in), M, vp(ko), vp(vo), stream), rast._h)
    run()
    torch.cuda.synchronize()
    assert torch.equal(ko.view(torch.int64), keys_ref.view(torch.int64)) and torch.equal(vo.view(torch.int32), vals_ref.view(torch.int32))
    ms = timeit(run)
    res["results"]["gsr_sort_pairs"] = {"ms": round(ms, 4), "pairs_per_s": round(M / ms * 1e3 / 1e9, 2),
                                        "what": "onesweep over the 40 significant bits (5 passes over M) + histogram pass"}
    # the pipeline's own figure
    del rast
    rast = GaussianRasterizer(width=W, height=H, mode=mode)
    rast.profile(True)
    acc = {}
    for _ in range(10):
        P.gpu_forward(rast, dev, cam, deg)
        torch.cuda.synchronize()
        for k, v in rast.stage_times_ms().items():
            acc[k] = acc.get(k, 0.0) + v / 10
    ms = acc["presort"] + acc["sort"]
    res["results"]["pipeline"] = {"ms": round(ms, 4), "pairs_per_s": round(M / ms * 1e3 / 1e9, 2),
                                  "presort_ms": round(acc["presort"], 4), "sort_ms": round(acc["sort"], 4),
                                  "duplicate_ms": round(acc["duplicate"], 4),
                                  "what": "depth pre-sort of the N Gaussians (4 passes over N) + 2 tile-digit passes over M"}
    res["speedup_over_cub"] = {k: round(res["results"]["cub"]["ms"] / v["ms"], 2) for k, v in res["results"].items() if k != "cub"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "cub_yardstick.json"), "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
