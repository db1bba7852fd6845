#!/usr/bin/env python
"""Sweep SpMV launch/layout options on one GPU and print achieved algorithmic GB/s.

  python tools/sweep_spmv.py [--cells 185] [--fe 2] [--reps 10] [--mode grid|variants|all]

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

DEFAULTS = {"spmv.kernel": 0, "spmv.exact": 1, "spmv.threads": 256, "spmv.ctas_per_sm": 0,
            "spmv.slices_per_cta": 64, "spmv.variant": 0, "spmv.prefetch": 0}


def set_opts(**kw):
    for k, v in DEFAULTS.items():
        capi.set_option(k, kw.get(k.split(".")[1], v))


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
    ap.add_argument("--mode", default="all")
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
        if "ms" in kw and "gbs" not in kw:
            kw["gbs"] = round(bytes_alg / (kw["ms"] / 1e3) / 1e9, 1)
            kw["frac"] = round(kw["gbs"] / peak, 4)
        print(json.dumps(kw), flush=True)
        out.write(json.dumps(kw) + "\n")
        out.flush()

    for numbering, sigma in [(1, 4096), (0, 4096)]:
        capi.set_option("sell.sigma", sigma)
        t0 = time.perf_counter()
        A = capi.Matrix.brick(args.fe, args.cells, numbering)
        capi.synchronize()
        build_s = time.perf_counter() - t0
        info = A.info()
        x, y = capi.Vector(n), capi.Vector(n)
        x.fill_hash()
        base = dict(numbering=numbering, sigma=sigma)
        report(event="matrix", n=n, nnz=nnz, padded=info["padded_nnz"] / nnz, build_s=build_s, **base)
        if args.mode in ("all", "grid"):
            for threads, spc in itertools.product([128, 256, 512], [4, 8, 16, 32, 64, 128, 256, 1024]):
                if spc * 32 < threads:
                    continue
                set_opts(threads=threads, slices_per_cta=spc)
                report(kernel="sell", threads=threads, spc=spc, ms=time_spmv(A, x, y, args.reps), **base)
        if args.mode in ("all", "variants"):
            for variant, prefetch, threads, spc in itertools.product([0, 1, 2, 3, 4, 5], [0, 1], [128, 256], [16, 64]):
                set_opts(threads=threads, slices_per_cta=spc, variant=variant, prefetch=prefetch)
                report(kernel="sell", variant=variant, prefetch=prefetch, threads=threads, spc=spc,
                       ms=time_spmv(A, x, y, args.reps), **base)
            set_opts(exact=0)
            report(kernel="sell", exact=0, ms=time_spmv(A, x, y, args.reps), **base)
        set_opts()
        if numbering == 1 and args.mode in ("all", "pcg"):
            report(kernel="sell_transposed", ms=time_spmv(A, x, y, args.reps, transposed=True), **base)
            b = capi.Vector(n)
            A.spmv_dev(x, b)
            for wide in (1, 0):
                capi.set_option("vec.wide", wide)
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
                pm = {k: v[0] / max(v[1], 1) for k, v in parts.items()}
                report(kernel="pcg_iteration", wide=wide, ms_per_iter=ms / 100, it_per_s=100 / (ms / 1e3),
                       it_gbs=it_bytes / (ms / 100 / 1e3) / 1e9, parts_ms=pm,
                       vec_update_gbs=7 * 8.0 * n / (pm["pcg.update"] / 1e3) / 1e9,
                       vec_direction_gbs=4 * 8.0 * n / (pm["pcg.direction"] / 1e3) / 1e9, **base)
            capi.set_option("vec.wide", 1)
        A.close()
        del x, y


if __name__ == "__main__":
    main()
