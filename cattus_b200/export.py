"""state_dict -> .cb2 weight blob: the exporter a `cuda-b200` engine adds next to the reference's
`_export_model_impl` cases (training/cattus_train/self_play.py:85-150; suffix table :153-162 -> ".cb2").

The blob carries BatchNorm folded into the preceding bias-free convolution, in fp32 and in PyTorch tensor layouts;
all device layouts (bf16 casts, K-major packing, NCHW->NHWC permutation of the FC columns) are produced by the
C++ engine at load time, so the file format is independent of the kernels.

Fold (eval-mode BN, eps = 1e-5, training/cattus_train/net_utils.py:14,30,33 -- affine only where the reference has
`affine=True`):   scale = gamma / sqrt(var + eps);   w' = w * scale[oc];   b' = beta - mean * scale.

Layout: 64-byte header (u32 LE: magic "CB2\\0", version 1, game, S, C_in, M, F, R, VH, PH, hidden=128, architecture tag 0, 4 x 0), then f32 LE:
  stem w[F,C,3,3] b[F] | R x (conv1 w[F,F,3,3] b[F], conv2 w b) | value conv w[VH,F] b[VH] | value fc1 w[128,VH*S*S] b[128]
  | value fc2 w[128] b[1] | policy conv w[PH,F] b[PH] | policy fc w[M,PH*S*S] b[M]
"""
from __future__ import annotations

import struct
from typing import Mapping

import numpy as np

from ._lib import GAME_IDS

MAGIC = 0x00324243
BN_EPS = 1e-5
VALUE_HIDDEN = 128


def _np(v) -> np.ndarray:
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def _fold(sd, conv_key: str, bn_prefix: str, affine: bool):
    w = _np(sd[conv_key]).astype(np.float64)
    mean = _np(sd[f"{bn_prefix}.running_mean"]).astype(np.float64)
    var = _np(sd[f"{bn_prefix}.running_var"]).astype(np.float64)
    gamma = _np(sd[f"{bn_prefix}.weight"]).astype(np.float64) if affine else np.ones_like(mean)
    beta = _np(sd[f"{bn_prefix}.bias"]).astype(np.float64) if affine else np.zeros_like(mean)
    scale = gamma / np.sqrt(var + BN_EPS)
    return (w * scale[:, None, None, None]).astype(np.float32), (beta - mean * scale).astype(np.float32)


def infer_dims(sd: Mapping[str, object], game: str) -> dict:
    """Architecture from the tensor shapes of a ConvNetV1 state_dict (net_utils.py:45-89)."""
    stem = _np(sd["_conv1._conv.weight"])
    f, c_in = stem.shape[0], stem.shape[1]
    r = 0
    while f"_residual_blocks.{r}._conv1.weight" in sd:
        r += 1
    vh = _np(sd["_value_head.0._conv.weight"]).shape[0]
    ph = _np(sd["_policy_head.0._conv.weight"]).shape[0]
    s2 = _np(sd["_value_head.2.weight"]).shape[1] // vh
    s = int(round(s2 ** 0.5))
    assert s * s == s2, "value head input is not VH * S * S"
    moves = _np(sd["_policy_head.2.weight"]).shape[0]
    return dict(game=game, board_size=s, planes=c_in, moves=moves, filters=f, blocks=r, value_channels=vh, policy_channels=ph)


PLANES = {"ttt": 3, "hex": 3, "chess": 18}  # position_to_planes: ttt/net.rs:14-24, hex/net.rs:14-24, chess/net/mod.rs:19-60


def export_simple_blob(sd: Mapping[str, object], game: str) -> bytes:
    """SimpleTwoHeadedModel (net_utils.py:92-121): header with architecture tag 1 and `hidden` = C * S * S, then f32 LE
    dense1 w[n,n] b[n] | dense2 w[n,n] b[n] | value w[n] b[1] | policy w[M,n] b[M]."""
    n = _np(sd["_dense1.weight"]).shape[0]
    planes = PLANES[game]
    s = int(round((n / planes) ** 0.5))
    assert planes * s * s == n, "SimpleTwoHeadedModel width is not planes * S * S for this game"
    moves = _np(sd["_policy_head.weight"]).shape[0]
    parts = [np.ascontiguousarray(_np(sd[k]), dtype="<f4").reshape(-1).tobytes()
             for k in ("_dense1.weight", "_dense1.bias", "_dense2.weight", "_dense2.bias", "_value_head.weight", "_value_head.bias",
                       "_policy_head.weight", "_policy_head.bias")]
    header = struct.pack("<16I", MAGIC, 1, GAME_IDS[game], s, planes, moves, 0, 0, 0, 0, n, 1, 0, 0, 0, 0)
    return header + b"".join(parts)


def export_blob(sd: Mapping[str, object], game: str) -> bytes:
    if "_dense1.weight" in sd:
        return export_simple_blob(sd, game)
    d = infer_dims(sd, game)
    parts = []

    def add(a):
        parts.append(np.ascontiguousarray(a, dtype="<f4").tobytes())

    w, b = _fold(sd, "_conv1._conv.weight", "_conv1._bn", True)
    add(w), add(b)
    for i in range(d["blocks"]):
        p = f"_residual_blocks.{i}"
        w, b = _fold(sd, f"{p}._conv1.weight", f"{p}._bn1", False)
        add(w), add(b)
        w, b = _fold(sd, f"{p}._conv2.weight", f"{p}._bn2", True)
        add(w), add(b)
    w, b = _fold(sd, "_value_head.0._conv.weight", "_value_head.0._bn", False)
    add(w), add(b)
    assert _np(sd["_value_head.2.weight"]).shape[0] == VALUE_HIDDEN
    add(_np(sd["_value_head.2.weight"])), add(_np(sd["_value_head.2.bias"]))
    add(_np(sd["_value_head.4.weight"]).reshape(-1)), add(_np(sd["_value_head.4.bias"]).reshape(-1))
    w, b = _fold(sd, "_policy_head.0._conv.weight", "_policy_head.0._bn", False)
    add(w), add(b)
    add(_np(sd["_policy_head.2.weight"])), add(_np(sd["_policy_head.2.bias"]))
    header = struct.pack("<16I", MAGIC, 1, GAME_IDS[game], d["board_size"], d["planes"], d["moves"], d["filters"], d["blocks"],
                         d["value_channels"], d["policy_channels"], VALUE_HIDDEN, 0, 0, 0, 0, 0)
    return header + b"".join(parts)


def export_model(sd: Mapping[str, object], game: str, path) -> None:
    """Counterpart of `export_model(model, path, inference_cfg, input_shape)` (self_play.py:70-82) for engine cuda-b200."""
    with open(path, "wb") as f:
        f.write(export_blob(sd, game))
