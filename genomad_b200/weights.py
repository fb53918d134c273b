"""
Weight loading for the IGLOO1D classifier.

The reference loads ``genomad/data/nn_classifier.h5`` with Keras' legacy-H5 loader, which maps
weights to layers BY ORDER (reference genomad/modules/nn_classification.py:309-310,
genomad/_paths.py:20-22).  This module reads either that very file (through genomad_b200.h5lite,
no h5py) or the flat ``.npz`` exported from it by tools/export_weights.py, checks every shape, and
returns a dict of numpy arrays in Keras layouts keyed by short names.

Layer order in the file (root attr ``layer_names`` / group ``model`` attr ``weight_names``):
  conv1d, conv1d_1, conv1d_2, igloo1d_kernel, igloo1d_kernel_1, dense, batch_normalization,
  then dense_1, batch_normalization_1, dense_2 -- i.e. creation order in igloo.py:45-82 / model.py:28-44,
  so igloo1d_kernel sits on conv #1's output and igloo1d_kernel_1 on conv #3's.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict

import numpy as np

DEFAULT_NPZ = Path(__file__).resolve().parent / "data" / "nn_classifier.npz"

_ENC = "/model/"
KEYS = {
    "c1w": (_ENC + "conv1d/kernel:0", (6, 257, 128)), "c1b": (_ENC + "conv1d/bias:0", (128,)),
    "c2w": (_ENC + "conv1d_1/kernel:0", (6, 128, 128)), "c2b": (_ENC + "conv1d_1/bias:0", (128,)),
    "c3w": (_ENC + "conv1d_2/kernel:0", (6, 128, 128)), "c3b": (_ENC + "conv1d_2/bias:0", (128,)),
    "d0w": (_ENC + "dense/kernel:0", (256, 512)), "d0b": (_ENC + "dense/bias:0", (512,)),
    "bn0g": (_ENC + "batch_normalization/gamma:0", (512,)), "bn0b": (_ENC + "batch_normalization/beta:0", (512,)),
    "bn0m": (_ENC + "batch_normalization/moving_mean:0", (512,)),
    "bn0v": (_ENC + "batch_normalization/moving_variance:0", (512,)),
    "d1w": ("/dense_1/dense_1/kernel:0", (512, 512)), "d1b": ("/dense_1/dense_1/bias:0", (512,)),
    "bn1g": ("/batch_normalization_1/batch_normalization_1/gamma:0", (512,)),
    "bn1b": ("/batch_normalization_1/batch_normalization_1/beta:0", (512,)),
    "bn1m": ("/batch_normalization_1/batch_normalization_1/moving_mean:0", (512,)),
    "bn1v": ("/batch_normalization_1/batch_normalization_1/moving_variance:0", (512,)),
    "d2w": ("/dense_2/dense_2/kernel:0", (512, 3)), "d2b": ("/dense_2/dense_2/bias:0", (3,)),
}
for _s, _g in ((0, "igloo1d_kernel"), (1, "igloo1d_kernel_1")):
    KEYS[f"ig{_s}_w_mult"] = (f"{_ENC}{_g}/w_mult:0", (1, 2100, 4, 128))
    KEYS[f"ig{_s}_w_summer"] = (f"{_ENC}{_g}/w_summer:0", (1, 512, 1))
    KEYS[f"ig{_s}_w_bias"] = (f"{_ENC}{_g}/w_bias:0", (1, 2100))
    KEYS[f"ig{_s}_w_qk"] = (f"{_ENC}{_g}/w_qk:0", (2100, 749))
    KEYS[f"ig{_s}_w_v"] = (f"{_ENC}{_g}/w_v:0", (1, 128, 128))
    KEYS[f"ig{_s}_random_patches"] = (f"{_ENC}{_g}/random_patches:0", (2100, 4, 1))

EXPECTED_ORDER = ["conv1d", "conv1d_1", "conv1d_2", "igloo1d_kernel", "igloo1d_kernel_1", "dense",
                  "batch_normalization"]


def _validate(raw: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    out = {}
    for short, (path, shape) in KEYS.items():
        if path not in raw:
            raise KeyError(f"weight {path} missing")
        a = np.asarray(raw[path])
        if tuple(a.shape) != shape:
            raise ValueError(f"{path}: shape {a.shape}, expected {shape}")
        want = np.int32 if short.endswith("random_patches") else np.float32
        if a.dtype != want:
            raise ValueError(f"{path}: dtype {a.dtype}, expected {np.dtype(want)}")
        out[short] = np.ascontiguousarray(a)
    for s in (0, 1):
        p = out[f"ig{s}_random_patches"]
        if p.min() < 0 or p.max() >= 5997:
            raise ValueError("patch index out of range")
    return out


def load_weights(path=None) -> Dict[str, np.ndarray]:
    """Load from ``.npz`` (default: the copy shipped in genomad_b200/data) or from a Keras legacy ``.h5``."""
    p = Path(path) if path else DEFAULT_NPZ
    if p.suffix == ".npz":
        z = np.load(p)
        raw = {k: z[k] for k in z.files if k.startswith("/")}
        order = [str(x) for x in z["__weight_order__"]] if "__weight_order__" in z.files else None
    else:
        from .h5lite import H5File
        f = H5File(p)
        raw = dict(f.datasets)
        order = list(f.attrs["/"].get("layer_names") or []) + ["|"] + list(f.attrs.get("/model", {}).get("weight_names") or [])
    if order:
        enc = order[order.index("|") + 1:] if "|" in order else []
        seen = []
        for name in enc:
            layer = name.split("/")[0]
            if layer not in seen:
                seen.append(layer)
        if seen and seen != EXPECTED_ORDER:
            raise ValueError(f"unexpected encoder layer order {seen}; weights are matched to layers by order")
    return _validate(raw)


def to_c_struct(w: Dict[str, np.ndarray], Weights, IglooW, BnW):
    """Fill the ctypes mirror of ``gnm_weights`` (include/gnm.h) with host pointers into ``w``."""
    def ptr(k):
        return w[k].ctypes.data_as(C.c_void_p)
    cw = Weights()
    cw.conv1_kernel, cw.conv1_bias = ptr("c1w"), ptr("c1b")
    cw.conv2_kernel, cw.conv2_bias = ptr("c2w"), ptr("c2b")
    cw.conv3_kernel, cw.conv3_bias = ptr("c3w"), ptr("c3b")
    for s in (0, 1):
        g = IglooW()
        g.w_mult, g.w_summer, g.w_bias = ptr(f"ig{s}_w_mult"), ptr(f"ig{s}_w_summer"), ptr(f"ig{s}_w_bias")
        g.w_qk, g.w_v, g.patches = ptr(f"ig{s}_w_qk"), ptr(f"ig{s}_w_v"), ptr(f"ig{s}_random_patches")
        cw.igloo[s] = g
    cw.dense0_kernel, cw.dense0_bias = ptr("d0w"), ptr("d0b")
    cw.bn0 = BnW(ptr("bn0g"), ptr("bn0b"), ptr("bn0m"), ptr("bn0v"))
    cw.dense1_kernel, cw.dense1_bias = ptr("d1w"), ptr("d1b")
    cw.bn1 = BnW(ptr("bn1g"), ptr("bn1b"), ptr("bn1m"), ptr("bn1v"))
    cw.dense2_kernel, cw.dense2_bias = ptr("d2w"), ptr("d2b")
    return cw
