#!/usr/bin/env python
"""Roofline characterisation sweep of the H.X hot path (BASELINE.json configs[4]): FE order x block width x
enrichment DoFs per cell on ONE B200, plus per-GPU shapes of configs[2] (order 6, B=128) and configs[3] (order 5,
B=1024 in column batches).  Every point goes through the C ABI, is checked with a size-independent property
(symmetry <HX,Z> = <X,HZ> on unconstrained rows, tolerance 1e-11) and reports the cell kernel against BOTH
rooflines: HBM (algorithmic bytes) and the FP64 tensor pipe (2*B*S2 flops vs the measured DMMA peak).

    python tools/sweep.py [--quick] [--out gpurun_out/sweep.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def mesh_for(p, B, n_enr, h_cell_gb, vec_gb):
    """cube of cells sized so the cell matrices take ~h_cell_gb and one block vector at most vec_gb"""
    n = (p + 1) ** 3 + n_enr
    c_h = h_cell_gb * 1e9 / (8.0 * n * n)
    c_v = vec_gb * 1e9 / (8.0 * B * p ** 3)
    c = max(27.0, min(c_h, c_v))
    m = max(3, int(round(c ** (1.0 / 3.0))))
    return (m, m, m)


def run_point(torch, capi, synth, p, B, n_enr, n_proj, peaks, h_cell_gb=1.5, vec_gb=1.5, degree=4, label=None):
    nc = mesh_for(p, B, n_enr, h_cell_gb, vec_gb)
    L = np.array(nc) * 0.8
    atoms = np.array([0.47 * L]) if (n_enr or n_proj) else None
    spec = synth.MeshSpec(ncell=nc, p=p, h=0.8, atoms=atoms, n_enr_per_atom=n_enr, enr_cutoff=1e3 if n_enr else 0.0,
                          n_proj_per_atom=n_proj, proj_cutoff=1.3 * 0.8, boundary="dirichlet")
    t0 = time.time()
    prob = synth.build_problem(spec)[0]
    t_build = time.time() - t0
    stream = torch.cuda.Stream()
    t0 = time.time()
    plan = capi.Plan(prob, max_block=B, stream=stream.cuda_stream)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, prob.diag_inv, prob.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    t_plan = time.time() - t0
    rows = prob.row_ids.astype(np.int64)
    rng = np.random.default_rng(11)
    X = rng.uniform(-0.5, 0.5, size=(prob.n_local, B))
    Z = rng.uniform(-0.5, 0.5, size=(prob.n_local, B))
    X[rows] = 0.0
    Z[rows] = 0.0

    def blk(host=None):
        t = torch.zeros(prob.n_local * B, dtype=torch.float64, device="cuda") if host is None else \
            torch.from_numpy(np.ascontiguousarray(host)).reshape(-1).cuda()
        o = type("Blk", (), {})()
        o.t, o.B, o.p = t, B, C.cast(t.data_ptr(), capi.f64p)
        return o

    with torch.cuda.stream(stream):
        dX, dZ, dY, dW = blk(X), blk(Z), blk(), blk()
        torch.cuda.synchronize()
        H.apply(dX, dY, True, False)
        H.apply(dZ, dW, True, False)
        plan.synchronize()
        n_own = prob.n_owned
        Xv, Zv = dX.t.view(-1, B)[:n_own], dZ.t.view(-1, B)[:n_own]
        a = (dY.t.view(-1, B)[:n_own] * Zv).sum(0)
        b = (Xv * dW.t.view(-1, B)[:n_own]).sum(0)
        sym = float(((a - b).abs().max() / a.abs().max()).item())
        finite = bool(torch.isfinite(dY.t).all().item())
        for _ in range(2):
            H.apply(dX, dY, True, False)
        plan.synchronize()
        reps = 10
        plan.enable_kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            H.apply(dX, dY, True, False)
        e1.record(stream)
        plan.synchronize()
        apply_ms = e0.elapsed_time(e1) / reps
        cell_ms, nl = plan.cell_kernel_time_ms()
        cell_ms /= max(nl, 1)
        plan.enable_kernel_timing(False)
        # filter: `degree` fused degrees
        dF = blk()
        capi.chebyshev_filter(H, minv, dX, dF, 2, -3.0, 1.0, 400.0)
        dX.t.copy_(dZ.t)
        plan.synchronize()
        e0.record(stream)
        capi.chebyshev_filter(H, minv, dX, dF, degree, -3.0, 1.0, 400.0)
        e1.record(stream)
        plan.synchronize()
        filt_ms = e0.elapsed_time(e1) / degree
    S2 = prob.S2 + (int(np.sum(prob.num_cell_proj.astype(np.int64) * prob.num_cell_dofs.astype(np.int64)))
                    if prob.num_cell_proj is not None else 0)
    alg = 8 * S2 + 16 * B * prob.n_local + 4 * prob.S + 12 * prob.col_vals.size + 16 * len(prob.row_ids)
    flops = 2.0 * B * S2
    t_hbm = alg / (peaks["hbm_gbs"] * 1e9) * 1e3
    t_dmma = flops / (peaks["dmma_tflops"] * 1e12) * 1e3
    res = {"label": label, "p": p, "B": B, "n_enr_per_cell": n_enr, "n_proj": n_proj, "cells": list(nc),
           "n_c_max": int(prob.num_cell_dofs.max()), "dofs": int(prob.n_owned), "h_cell_gb": 8 * S2 / 1e9,
           "apply_ms": apply_ms, "cell_kernel_ms": cell_ms, "filter_ms_per_degree": filt_ms,
           "gdofvec_per_s": prob.n_owned * B / (apply_ms * 1e-3) / 1e9,
           "filter_gdofvec_per_s": prob.n_owned * B / (filt_ms * 1e-3) / 1e9,
           "alg_gbs": alg / (cell_ms * 1e-3) / 1e9, "tflops": flops / (cell_ms * 1e-3) / 1e12,
           "frac_hbm": t_hbm / cell_ms, "frac_dmma": t_dmma / cell_ms,
           "bound": "hbm" if t_hbm >= t_dmma else "dmma", "roofline_frac": max(t_hbm, t_dmma) / cell_ms,
           "arith_intensity": flops / alg, "symmetry_err": sym, "finite": finite, "ok": bool(finite and sym < 1e-11),
           "host_build_s": t_build, "plan_s": t_plan}
    del dX, dZ, dY, dW, dF, H, minv, plan
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--points", default="", help="comma-separated p:B[:enr[:proj]] points instead of the full grid")
    args = ap.parse_args()
    import torch
    from dft_efe_b200 import capi, synth
    assert torch.cuda.is_available(), "sweep needs a CUDA device"
    capi.check(capi.lib().hx_set_device(0))
    micro = capi.microbench()
    peaks = {"dmma_tflops": micro["dmma_tflops"], "hbm_gbs": micro["copy_gbs"], "source": "hx_microbench on this box"}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"] = float(mp["hbm_gbs"])
        peaks["source"] = "hbm: MEASURED_PEAKS.json; dmma: hx_microbench on this box"
    except Exception:
        pass
    points = []
    if args.points:
        grid = []
        for tok in args.points.split(","):
            f = [int(x) for x in tok.split(":")]
            grid.append((f[0], f[1], f[2] if len(f) > 2 else 0, f[3] if len(f) > 3 else 0))
    elif args.quick:
        grid = [(4, 32, 0, 0), (6, 128, 0, 0), (2, 16, 0, 0), (4, 32, 16, 4)]
    else:
        grid = [(p, B, 0, 0) for p in (2, 3, 4, 5, 6, 7, 8) for B in (16, 32, 128, 512, 2048)]
        grid += [(4, 32, e, 4) for e in (8, 16, 32, 64)] + [(6, 128, e, 4) for e in (16, 64)]
    out = {"peaks": peaks, "points": points}
    for (p, B, e, npj) in grid:
        try:
            r = run_point(torch, capi, synth, p, B, e, npj, peaks)
        except Exception as ex:  # noqa: BLE001
            r = {"p": p, "B": B, "n_enr_per_cell": e, "error": str(ex)[:300], "ok": False}
        points.append(r)
        print(json.dumps(r), flush=True)
        json.dump(out, open(args.out, "w"), indent=1)
    # per-GPU shapes of the multi-GPU configs (one of 8 slabs), single GPU
    if not args.quick and not args.points:
        for label, p, B, e, hg, vg in (("C3/8: benzene-dimer-like slab, order 6, B=128", 6, 128, 8, 4.5, 1.1),
                                       ("C4/8 column batch: order 5, B=256 of 1024", 5, 256, 0, 6.0, 4.0)):
            try:
                r = run_point(torch, capi, synth, p, B, e, 4 if e else 0, peaks, h_cell_gb=hg, vec_gb=vg, label=label)
            except Exception as ex:  # noqa: BLE001
                r = {"label": label, "error": str(ex)[:300], "ok": False}
            points.append(r)
            print(json.dumps(r), flush=True)
            json.dump(out, open(args.out, "w"), indent=1)
    bad = [r for r in points if not r.get("ok")]
    print(f"{len(points)} points, {len(bad)} failed")


if __name__ == "__main__":
    main()
