#!/usr/bin/env python
"""Regenerates tests/golden/ref_hx_small.npz and ref_filter_small.npz: H.X (and a degree-9 Chebyshev filter, see below)
of a small full-feature problem (hanging nodes, Dirichlet rows,
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

    # ref_filter_small.npz: the same problem through the REFERENCE'S OWN ChebyshevFilter template
    # (linearAlgebra/ChebyshevFilter.t.cpp:39-134, compiled into oracle/_ref) driving the reference-assembled apply above;
    # the mass-lumped M^-1 apply (OEFEAtomBlockOverlapInvOpContextGLL, a deal.II-dependent class that cannot be compiled
    # here) is the oracle port.  Pins filter scalars, recurrence and operator order for the oracle and the CUDA path.
    import ctypes as C
    from oracle import oracle as orc
    W = orc.OracleWorld([p])
    deg, a0, a, b = 9, -3.0, 1.0, 60.0

    def cb(_user, op_id, xp, yp, n_, B_, ugx, ugy):
        Xa = np.ctypeslib.as_array(xp, shape=(n_, B_))
        Ya = np.ctypeslib.as_array(yp, shape=(n_, B_))
        if op_id == 0:
            ref.hx_apply_serial(p, Xa, cell_block=1, out=Ya)
        else:
            W.minv_apply([Xa], [Ya], bool(ugx), bool(ugy))

    xr, yr = X.copy(), np.zeros_like(X)
    ref.lib().ref_chebyshev_filter(ref.APPLY_CB(cb), None, orc._f64(xr), orc._f64(yr), C.c_uint32(p.n_local), C.c_uint32(B),
                                   C.c_uint32(deg), C.c_double(a0), C.c_double(a), C.c_double(b))
    np.savez_compressed(os.path.join(HERE, "ref_filter_small.npz"), p=p_order, nc=np.array(nc), B=B, degree=deg,
                        bounds=np.array([a0, a, b]), X=X, F=yr[:p.n_owned])
    print("wrote ref_filter_small.npz", yr.shape, float(np.abs(yr[:p.n_owned]).max()))


if __name__ == "__main__":
    main()
