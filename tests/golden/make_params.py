"""Generate the URS fixtures tests/golden/params_vesta_k{K}.npz with the oracle's restatement of
`Params::<vesta::Affine>::new(k)` (oracle/halo2.py: hash_to_curve("Halo2-Parameters") + EC inverse FFT).
In a Rust deployment this object comes from halo2_proofs itself (`Params::new` / `Params::read`); the bench and the
GPU tests load these arrays instead of recomputing them (k=12 takes ~10 s of host time).
Arrays: g, g_lagrange (n x 8 uint64, Montgomery affine x||y), w, u (8 uint64).
Run:  python tests/golden/make_params.py 5 11 12"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import halo2 as H

if __name__ == "__main__":
    for k in [int(a) for a in sys.argv[1:]] or [5, 11, 12]:
        p = H.Params.new(k, 0)
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"params_vesta_k{k}.npz")
        np.savez_compressed(out, g=p.g, g_lagrange=p.g_lagrange, w=p.w, u=p.u)
        print("wrote", out, os.path.getsize(out))
