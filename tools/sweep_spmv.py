#!/usr/bin/env python
"""Sweep SpMV launch/layout options on one GPU and print achieved algorithmic GB/s.

  python tools/sweep_spmv.py [--cells 185] [--fe 2] [--reps 10] [--quick]

Algorithmic bytes per launch = 12 nnz + 4 (n+1) + 16 n; time = CUDA events
on the library's stream around `reps` back-to-back launches (inputs far exceed L2)."""
import argparse
import itertools
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi  # noqa: E402


def time_spmv(A, x, y, reps, transposed=False):
    f = A.spmv_transpose_dev if transposed else A.spmv_dev
    for _ in range(2):
        f(x, y)
    capi.synchronize()
    capi.timer_start()
    for _ in range(reps):
        f(x, y)
    return capi.timer_stop() / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=185)
    ap.add_argument("--fe", type=int, default=2)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/sweep_spmv.jsonl")
    args = ap.parse_args()
    capi.init(0)
    n, nnz = capi.brick_sizes(args.fe, args.cells)
    bytes_alg = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    out = open(args.out, "a")
    peak = 6554.9
    try:
        peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        pass

    def report(**kw):
        kw["gbs"] = bytes_alg / (kw["ms"] / 1e3) / 1e9
        kw["frac"] = kw["gbs"] / peak
        print(json.dumps(kw), flush=True)
        out.write(json.dumps(kw) + "\n")
        out.flush()

    sigmas = [4096] if args.quick else [4096, 1024, 0]
    numberings = [1] if args.quick else [1, 0]
    for numbering, sigma in itertools.product(numberings, sigmas):
        if numbering == 0 and sigma != 4096:
            continue
        capi.set_option("sell.sigma", sigma)
        t0 = time.perf_counter()
        A = capi.Matrix.brick(args.fe, args.cells, numbering)
        capi.synchronize()
        build_s = time.perf_counter() - t0
        info = A.info()
        x, y = capi.Vector(n), capi.Vector(n)
        x.fill_hash()
        base = dict(n=n, nnz=nnz, numbering=numbering, sigma=sigma, padded=info["padded_nnz"] / nnz, build_s=build_s)
        capi.set_option("spmv.kernel", 0)
        for exact in ([1] if args.quick else [1, 0]):
            capi.set_option("spmv.exact", exact)
            for threads, ctas in itertools.product([256, 512, 1024], [1, 2, 4, 8, 32, 128]):
                if threads * ctas < 1024 or (threads * ctas > 2048 and ctas <= 8):
                    continue
                if exact == 0 and not (threads == 512 and ctas in (4, 32)):
                    continue
                capi.set_option("spmv.threads", threads)
                capi.set_option("spmv.ctas_per_sm", ctas)
                ms = time_spmv(A, x, y, args.reps)
                report(kernel="sell", exact=exact, threads=threads, ctas_per_sm=ctas, ms=ms, **base)
        capi.set_option("spmv.exact", 1)
        capi.set_option("spmv.threads", 512)
        capi.set_option("spmv.ctas_per_sm", 4)
        if sigma == 4096:
            capi.set_option("spmv.kernel", 1)
            for lanes in [4, 8, 16, 32]:
                capi.set_option("spmv.lanes", lanes)
                ms = time_spmv(A, x, y, args.reps)
                report(kernel="csr_vector", lanes=lanes, ms=ms, **base)
            capi.set_option("spmv.kernel", 0)
            if numbering == 1:
                ms = time_spmv(A, x, y, args.reps, transposed=True)
                report(kernel="sell_transposed", ms=ms, **base)
                # whole PCG iteration and its three kernels
                b = capi.Vector(n)
                A.spmv_dev(x, b)
                y.fill(0.0)
                A.pcg_dev(y, b, 20, 0.0, True)
                capi.set_option("timing", 1)
                capi.timing_reset()
                y.fill(0.0)
                capi.synchronize()
                capi.timer_start()
                A.pcg_dev(y, b, 100, 0.0, True)
                ms = capi.timer_stop()
                parts = {k: capi.timing(k) for k in ("spmv", "pcg.update", "pcg.direction", "pcg.setup")}
                capi.set_option("timing", 0)
                it_bytes = 12.0 * nnz + 4.0 * (n + 1) + 104.0 * n
                rec = dict(kernel="pcg_iteration", ms=ms / 100, it_per_s=100 / (ms / 1e3),
                           it_gbs=it_bytes / (ms / 100 / 1e3) / 1e9,
                           parts_ms={k: v[0] / max(v[1], 1) for k, v in parts.items()}, **base)
                rec["vec_update_gbs"] = 7 * 8.0 * n / (rec["parts_ms"]["pcg.update"] / 1e3) / 1e9
                rec["vec_direction_gbs"] = 4 * 8.0 * n / (rec["parts_ms"]["pcg.direction"] / 1e3) / 1e9
                print(json.dumps(rec), flush=True)
                out.write(json.dumps(rec) + "\n")
        A.close()
        del x, y


if __name__ == "__main__":
    main()
