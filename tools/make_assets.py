"""Convert the reference's example OBJ input into a compact binary fixture.

Run in the build container (the reference checkout does not exist on the GPU box):

    python tools/make_assets.py

teapot.obj (examples/teapot.rs:16) is parsed with rpt_b200.api.parse_obj -- the mirror
of src/io.rs:27-73 -- and stored as an (n, 18) float64 triangle array
(v1,v2,v3,n1,n2,n3 per row), the exact input Mesh::new receives in the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpt_b200.api import parse_obj  # noqa: E402

REF = os.environ.get("RPT_REFERENCE", "/root/reference")


def main():
    with open(os.path.join(REF, "examples", "teapot.obj")) as f:
        tris = parse_obj(f)
    out = os.path.join(ROOT, "rpt_b200", "assets", "teapot_tris.npz")
    np.savez_compressed(out, tris=tris)
    print(out, tris.shape, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
