#!/usr/bin/env python
"""Roofline characterisation of the subspace GEMMs (BASELINE north_star: "X^T H X and X^T X as
tensor-core GEMMs ... FP64 tensor-pipe utilisation against the FP64 tensor peak") and of the dense B x B steps, on ONE
B200, through the C ABI: for each block width B one CholeskyGramSchmidt + one RayleighRitz call on a random block with
the mass-lumped overlap operator (cheap apply, so the Gram GEMM / rotation GEMM / cuSOLVER phases dominate), timed by
the plan's phase trace (CUDA events on the plan's stream).

    python tools/subspace_bench.py [--cells 20] [--order 4] [--widths 32,128,256,512,1024] [--out gpurun_out/subspace.json]

Every point is verified by a size-independent property: after CholeskyGramSchmidt X^T M X = I (1e-10), after
RayleighRitz X^T M X = I still and the Ritz values ascend.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=20)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--widths", default="32,128,256,512,1024")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "subspace.json"))
    args = ap.parse_args()
    from dft_efe_b200 import capi, synth
    assert capi.device_count() >= 1, "needs a CUDA device: libhxb200 has no CPU fallback"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    micro = capi.microbench()
    dmma = micro["dmma_tflops"]
    nc = (args.cells,) * 3
    prob = synth.build_problem(synth.MeshSpec(ncell=nc, p=args.order, h=0.8, boundary="dirichlet"))[0]
    N = prob.n_owned
    rows = []
    for B in [int(w) for w in args.widths.split(",")]:
        plan = capi.Plan(prob, max_block=B)
        M = capi.DiagOp(plan, prob.diag, prob.enr_block, capi.DIAG_OEFE_MASS)
        rng = np.random.default_rng(B)
        X = rng.uniform(-0.5, 0.5, (prob.n_local, B))
        X[prob.row_ids.astype(np.int64)] = 0.0
        dX, dO, dV = plan.block(B, X), plan.block(B), plan.block(B)
        del X
        acc = {}
        for rep in range(args.reps + 1):
            dX2 = dX  # CholGS orthonormalises in place: re-orthonormalising an orthonormal block is the same work
            plan.trace(True)
            st = capi.cholesky_gram_schmidt(M, dX2, dO, B)
            assert st == 0, st
            w, st2 = capi.rayleigh_ritz(M, dO, dV, B, True)
            assert st2 == 0
            rep_ = plan.trace_report()
            plan.trace(False)
            if rep == 0:
                continue  # warm-up (cuSOLVER handle, workspace)
            for k, v in rep_.items():
                a = acc.setdefault(k, [0.0, 0])
                a[0] += v["ms"]
                a[1] += v["n"]
        ms = {k: v[0] / v[1] for k, v in acc.items()}
        # property check: M-orthonormal after both steps, Ritz values ascending
        S = capi.xtopx_device(M, dV, B).download()
        S = S + S.T - np.diag(np.diag(S))
        ortho_err = float(np.abs(S - np.eye(B)).max())
        assert ortho_err < 1e-10, ortho_err
        assert np.all(np.diff(w) >= -1e-12)
        tiles = (B + 63) // 64
        gram_flops = 2.0 * N * 64 * 64 * (tiles * (tiles + 1) / 2)  # tiles on and below the diagonal
        gram_bytes = 16.0 * N * B
        rot_flops_full = 2.0 * N * B * B
        row = {"B": B, "N": N, "ms": {k: round(v, 4) for k, v in ms.items()},
               "gram": {"ms": ms.get("gram"), "tflops": gram_flops / (ms["gram"] * 1e-3) / 1e12,
                        "frac_dmma": gram_flops / (ms["gram"] * 1e-3) / 1e12 / dmma,
                        "gbs": gram_bytes / (ms["gram"] * 1e-3) / 1e9, "frac_hbm": gram_bytes / (ms["gram"] * 1e-3) / 1e9 / hbm},
               "rotate": {"ms": ms.get("rotate"), "tflops": rot_flops_full / (ms["rotate"] * 1e-3) / 1e12,
                          "frac_dmma": rot_flops_full / (ms["rotate"] * 1e-3) / 1e12 / dmma,
                          "gbs": 16.0 * N * B / (ms["rotate"] * 1e-3) / 1e9},
               "rotate_lower": {"ms": ms.get("rotate-lower"),
                                "tflops": 0.5 * rot_flops_full * (1 + 1.0 / tiles) / (ms["rotate-lower"] * 1e-3) / 1e12},
               "dense_cholesky_ms": ms.get("dense-cholesky"), "dense_eig_ms": ms.get("dense-eig"),
               "orthonormality_error": ortho_err}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del dX, dO, dV, M, plan
    out = {"workload": f"{nc[0]}^3 cells, order {args.order}, {N} DoFs", "dmma_peak_tflops": dmma, "hbm_peak_gbs": hbm,
           "points": rows}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
