#!/usr/bin/env python
"""
Convert geNomad's Keras legacy-H5 weight file into the flat ``.npz`` this repo ships.

    python tools/export_weights.py [/root/reference/genomad/data/nn_classifier.h5] [out.npz]

The H5 is read with genomad_b200.h5lite (no h5py).  Keys of the npz are the HDF5
dataset paths (e.g. ``/model/conv1d/kernel:0``); values are the raw arrays, bit-exact.
Two extra entries record provenance: ``__sha256__`` of the source file and
``__weight_order__`` (the ``weight_names``/``layer_names`` attributes Keras uses to map
weights to layers by ORDER -- reference ``genomad/modules/nn_classification.py:310``).
The reference file is data, not source; it is needed on the GPU box where
``/root/reference`` does not exist.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from genomad_b200.h5lite import H5File  # noqa: E402

EXPECTED_SHA256 = "834bcb03aeb1ff484dc7c1f7c00fb951708a91ed03d8a233cd176390930761a1"


def main():
    src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/genomad/data/nn_classifier.h5")
    dst = Path(sys.argv[2] if len(sys.argv) > 2 else
               Path(__file__).resolve().parents[1] / "genomad_b200" / "data" / "nn_classifier.npz")
    sha = hashlib.sha256(src.read_bytes()).hexdigest()
    if sha != EXPECTED_SHA256:
        print(f"warning: sha256 {sha} differs from the surveyed file {EXPECTED_SHA256}", file=sys.stderr)
    f = H5File(src)
    order = list(f.attrs["/"]["layer_names"]) + ["|"] + list(f.attrs["/model"]["weight_names"])
    arrays = {k: np.ascontiguousarray(v) for k, v in f.datasets.items()}
    arrays["__sha256__"] = np.array(sha)
    arrays["__weight_order__"] = np.array(order)
    np.savez(dst, **arrays)
    print(f"wrote {dst} ({dst.stat().st_size} bytes, {len(f.datasets)} datasets)")


if __name__ == "__main__":
    main()
