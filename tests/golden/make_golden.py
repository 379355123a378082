#!/usr/bin/env python
"""Regenerates tests/golden/ref_hx_small.npz: H.X of a small full-feature problem (hanging nodes, Dirichlet rows,
enrichment, nonlocal projectors) computed by KohnShamOperatorContextFE::apply ASSEMBLED FROM THE REFERENCE'S OWN
COMPILED ROUTINES (oracle/_ref/libdftefe_ref.so, ref_hx_apply_serial in oracle/ref_shim_cellwise.cpp).  Needs
/root/reference (dev container); the fixture travels so that the oracle and the GPU path can be pinned against it
anywhere.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from dft_efe_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402
from tests.test_oracle import small_spec  # noqa: E402


def main():
    assert ref.build(), "oracle/_ref could not be built (needs /root/reference)"
    p_order, nc, B = 3, (4, 4, 4), 6
    p = synth.build_problem(small_spec(1, p=p_order, nc=nc))[0]
    X = synth.make_block(p, B)
    Xr = X.copy()
    Y = ref.hx_apply_serial(p, Xr, cell_block=3)
    np.savez_compressed(os.path.join(HERE, "ref_hx_small.npz"), p=p_order, nc=np.array(nc), B=B, X=X, X_after=Xr, Y=Y)
    print("wrote ref_hx_small.npz", Y.shape, float(np.abs(Y).max()))


if __name__ == "__main__":
    main()
