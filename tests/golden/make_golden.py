"""Regenerates tests/golden/*.npz.

These fixtures are OUTPUTS OF THE ORACLE (oracle/radlite_oracle.c), not of the Fortran reference:
the reference cannot be compiled in this image and ships no expected outputs,
so parity stays "unpinned".  They freeze the oracle's behaviour so that (a) an accidental change of
the oracle shows up on CPU and (b) the CUDA path is also compared against committed numbers.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import golden_cases as cases  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from radlite_b200 import synth  # noqa: E402


def main():
    for name, m in cases():
        o = Oracle()
        o.load_model(m)
        out = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
        c = o.counters()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), flux=out["flux"], velo=out["velo"],
                            tau_center=out["tau_center"], maserflag=out["maserflag"],
                            image_ring=out["image"][:, ::7, :, :], cmask_ring=out["cmask"][:, ::7].astype(np.int8),
                            counters=np.array([c["R"], c["E"], c["S"]]))
        print(name, out["flux"].shape, c)


if __name__ == "__main__":
    main()
