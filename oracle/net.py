"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the floating-point half of the path: the ConvNetV1 forward pass (and
the reference's second model type, SimpleTwoHeadedModel, net_utils.py:92-121).

Restates `training/cattus_train/net_utils.py:4-89` (ConvBlock :4-20, ResidualBlock :23-42, ConvNetV1 :45-89)
as explicit fp32 torch-functional calls over a plain `{state_dict key: numpy array}` mapping, so that it runs
on the GPU box where /root/reference does not exist.  It is what the reference's torch-py engine computes
(`engine/src/net/model.rs:68-84`: torch module, `from_numpy` in, `.numpy()` out, `no_grad`) and what
`training/tests/test_net_output.py:28-33` accepts as ground truth for the Rust engines.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs import this.

Parity pinning: `oracle/gen_golden.py` (run in the build container) loads the REAL `ConvNetV1` from
/root/reference by file path, loads the state_dict produced by `make_state_dict` below into it, and stores its
outputs in `tests/golden/net_ref_*.npz`; `tests/test_oracle_golden.py` replays the same seeds through
`convnet_forward` and requires agreement to fp32 round-off.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, Tuple

import numpy as np

BN_EPS = 1e-5  # torch BatchNorm2d default, not overridden by the reference (net_utils.py:14,30,33)
VALUE_HIDDEN = 128  # net_utils.py:71-73


@dataclass(frozen=True)
class NetConfig:
    board_size: int  # S
    planes: int  # C_in
    moves: int  # M
    filters: int  # residual_filter_num F
    blocks: int  # residual_block_num R
    value_channels: int  # VH
    policy_channels: int  # PH
    game: str = "hex"  # "ttt" | "hex" | "chess"
    arch: str = "convnet"  # "convnet" = ConvNetV1 (net_utils.py:45-89) | "simple" = SimpleTwoHeadedModel (:92-121; F, R, VH, PH unused)

    def to_dict(self):
        return asdict(self)

    @property
    def features(self) -> int:
        """SimpleTwoHeadedModel's width: C * H * W (net_utils.py:97-99)."""
        return self.planes * self.board_size ** 2

    @property
    def flops_per_position(self) -> int:
        """2*MAC, unpadded."""
        s2 = self.board_size ** 2
        if self.arch == "simple":
            n = self.features
            return 2 * (2 * n * n) + 2 * n + 2 * n * self.moves
        f, r = self.filters, self.blocks
        return (2 * 9 * self.planes * f * s2 + r * 2 * (2 * 9 * f * f * s2) + 2 * f * self.value_channels * s2
                + 2 * self.value_channels * s2 * VALUE_HIDDEN + 2 * VALUE_HIDDEN + 2 * f * self.policy_channels * s2
                + 2 * self.policy_channels * s2 * self.moves)

    @property
    def trunk_flops_per_position(self) -> int:
        """The R residual blocks only (the figure BASELINE.md quotes for the 50%-of-peak target)."""
        s2 = self.board_size ** 2
        return self.blocks * 2 * (2 * 9 * self.filters * self.filters * s2)


# The configurations BASELINE.json / training/config/*.yaml name.
CONFIGS: Dict[str, NetConfig] = {
    "ttt": NetConfig(3, 3, 9, 16, 3, 8, 8, "ttt"),
    "hex4": NetConfig(4, 3, 16, 16, 7, 16, 16, "hex"),
    "hex5": NetConfig(5, 3, 25, 16, 7, 16, 16, "hex"),
    "hex7": NetConfig(7, 3, 49, 16, 7, 16, 16, "hex"),
    "hex9": NetConfig(9, 3, 81, 16, 7, 16, 16, "hex"),
    "hex11": NetConfig(11, 3, 121, 16, 7, 16, 16, "hex"),
    "chess_dev": NetConfig(8, 18, 1880, 16, 7, 8, 8, "chess"),
    "chess10x128": NetConfig(8, 18, 1880, 128, 10, 32, 32, "chess"),
    # the tiny nets the reference's own parity tests build (test_net_output.py:110-117, test_convnetv1.py:21-24)
    "ttt_1x1": NetConfig(3, 3, 9, 1, 1, 1, 1, "ttt"),
    "hex11_1x1": NetConfig(11, 3, 121, 1, 1, 1, 1, "hex"),
    "chess_1x1": NetConfig(8, 18, 1880, 1, 1, 1, 1, "chess"),
    "hex5_2x2": NetConfig(5, 3, 25, 2, 2, 4, 4, "hex"),
    "chess_2x128": NetConfig(8, 18, 1880, 128, 2, 32, 32, "chess"),
    # widths and depths the reference's config recommends (training/config/chess_dev.yaml:30-37: 32-256 filters, 7-39 blocks)
    # that neither whole-trunk kernel covers: they take the per-layer tensor-core path
    "chess_4x64": NetConfig(8, 18, 1880, 64, 4, 8, 8, "chess"),
    "chess_4x256": NetConfig(8, 18, 1880, 256, 4, 16, 16, "chess"),
    "hex7_4x32": NetConfig(7, 3, 49, 32, 4, 16, 16, "hex"),
    "hex11_2x128": NetConfig(11, 3, 121, 128, 2, 16, 16, "hex"),
    "chess20x256": NetConfig(8, 18, 1880, 256, 20, 32, 32, "chess"),  # the top of the recommended range (bench workload only)
    # SimpleTwoHeadedModel (net_utils.py:92-121; training/tests/test_simple_two_headed.py)
    "ttt_simple": NetConfig(3, 3, 9, 0, 0, 0, 0, "ttt", "simple"),
    "hex5_simple": NetConfig(5, 3, 25, 0, 0, 0, 0, "hex", "simple"),
    "hex11_simple": NetConfig(11, 3, 121, 0, 0, 0, 0, "hex", "simple"),
    "chess_simple": NetConfig(8, 18, 1880, 0, 0, 0, 0, "chess", "simple"),
    # depth ladder for the fused-trunk parity tests (stem only, one block)
    "chess_0x128": NetConfig(8, 18, 1880, 128, 0, 32, 32, "chess"),
    "chess_1x128": NetConfig(8, 18, 1880, 128, 1, 32, 32, "chess"),
}


def state_dict_spec(cfg: NetConfig):
    """(key, shape, kind) for every tensor of ConvNetV1's state_dict, in module order.
    kind: conv | bn_w | bn_b | bn_mean | bn_var | bn_nbt | fc_w | fc_b."""
    s2 = cfg.board_size ** 2
    if cfg.arch == "simple":  # net_utils.py:101-110
        n = cfg.features
        return [("_dense1.weight", (n, n), "fc_w"), ("_dense1.bias", (n,), "fc_b"), ("_dense2.weight", (n, n), "fc_w"), ("_dense2.bias", (n,), "fc_b"),
                ("_value_head.weight", (1, n), "fc_w"), ("_value_head.bias", (1,), "fc_b"),
                ("_policy_head.weight", (cfg.moves, n), "fc_w"), ("_policy_head.bias", (cfg.moves,), "fc_b")]
    f = cfg.filters
    spec = [("_conv1._conv.weight", (f, cfg.planes, 3, 3), "conv")]
    spec += _bn("_conv1._bn", f, affine=True)
    for i in range(cfg.blocks):
        p = f"_residual_blocks.{i}"
        spec.append((f"{p}._conv1.weight", (f, f, 3, 3), "conv"))
        spec += _bn(f"{p}._bn1", f, affine=False)
        spec.append((f"{p}._conv2.weight", (f, f, 3, 3), "conv"))
        spec += _bn(f"{p}._bn2", f, affine=True)
    vh, ph = cfg.value_channels, cfg.policy_channels
    spec.append(("_value_head.0._conv.weight", (vh, f, 1, 1), "conv"))
    spec += _bn("_value_head.0._bn", vh, affine=False)
    spec += [("_value_head.2.weight", (VALUE_HIDDEN, vh * s2), "fc_w"), ("_value_head.2.bias", (VALUE_HIDDEN,), "fc_b"),
             ("_value_head.4.weight", (1, VALUE_HIDDEN), "fc_w"), ("_value_head.4.bias", (1,), "fc_b")]
    spec.append(("_policy_head.0._conv.weight", (ph, f, 1, 1), "conv"))
    spec += _bn("_policy_head.0._bn", ph, affine=False)
    spec += [("_policy_head.2.weight", (cfg.moves, ph * s2), "fc_w"), ("_policy_head.2.bias", (cfg.moves,), "fc_b")]
    return spec


def _bn(prefix: str, c: int, affine: bool):
    out = []
    if affine:
        out += [(f"{prefix}.weight", (c,), "bn_w"), (f"{prefix}.bias", (c,), "bn_b")]
    out += [(f"{prefix}.running_mean", (c,), "bn_mean"), (f"{prefix}.running_var", (c,), "bn_var"),
            (f"{prefix}.num_batches_tracked", (), "bn_nbt")]
    return out


def make_state_dict(cfg: NetConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Deterministic random-init weights (numpy PCG64, stable across machines) with randomised BN statistics
    (mean ~ N(0,0.1), var ~ U(0.5,1.5)) so that the BN fold is exercised.  Convs use
    N(0, 1/fan_in) (keeps logits O(1) through 10 blocks, like a trained net), linears U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default range."""
    rng = np.random.default_rng(0xCA77 + seed)
    sd: Dict[str, np.ndarray] = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            a = rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)
        elif kind == "bn_w":
            a = rng.uniform(0.5, 1.5, shape)
        elif kind in ("bn_b", "bn_mean"):
            a = rng.standard_normal(shape) * 0.1
        elif kind == "bn_var":
            a = rng.uniform(0.5, 1.5, shape)
        elif kind == "bn_nbt":
            sd[key] = np.array(1, dtype=np.int64)
            continue
        elif kind == "fc_w":
            bound = 1.0 / np.sqrt(shape[1])
            a = rng.uniform(-bound, bound, shape)
        elif kind == "fc_b":
            a = rng.uniform(-0.1, 0.1, shape)
        else:
            raise AssertionError(kind)
        sd[key] = a.astype(np.float32)
    return sd


def convnet_forward(sd: Dict[str, np.ndarray], cfg: NetConfig, x: np.ndarray, threads: int | None = None
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """ConvNetV1.forward in eval mode (net_utils.py:84-89): x [B,C,S,S] f32 -> (policy logits [B,M], value [B,1])."""
    import torch

    if threads is not None:
        torch.set_num_threads(threads)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}

    with torch.no_grad():
        p_, v = _forward_torch(t, cfg, torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))
    return p_.numpy(), v.numpy()


class TorchCpuModel:
    """The reference's torch-py engine restated (engine/src/net/model.rs:68-84): a traced module on CPU,
    `from_numpy` in, `.numpy()` out, under `no_grad`.  Used only as the timed CPU baseline."""

    def __init__(self, sd: Dict[str, np.ndarray], cfg: NetConfig, batch: int, threads: int):
        import torch

        torch.set_num_threads(threads)
        self.torch = torch
        self.cfg = cfg
        self.sd = sd

        class _M(torch.nn.Module):
            def forward(inner, x):  # noqa: N805
                p, v = _forward_torch(self._t, cfg, x)
                return p, v

        self._t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
        ex = torch.zeros((batch, cfg.planes, cfg.board_size, cfg.board_size), dtype=torch.float32)
        with torch.no_grad():
            self.model = torch.jit.trace(_M().eval(), ex, check_trace=False)

    def run(self, x: np.ndarray):
        torch = self.torch
        with torch.no_grad():
            p, v = self.model(torch.from_numpy(x))
            return p.detach().cpu().numpy(), v.detach().cpu().numpy()


def _forward_torch(t, cfg: NetConfig, flow):
    """stem ConvBlock(bn_scale=True) net_utils.py:60,:4-20; ResidualBlock x R :23-42; value head :68-75
    (conv1x1 -> BN(no affine) -> ReLU -> flatten(NCHW) -> FC -> ReLU -> FC -> tanh); policy head :78-82 (raw logits)."""
    import torch
    import torch.nn.functional as F

    if cfg.arch == "simple":  # SimpleTwoHeadedModel.forward (net_utils.py:112-121)
        flow = flow.flatten(1)
        flow = F.relu(F.linear(flow, t["_dense1.weight"], t["_dense1.bias"]))
        flow = F.relu(F.linear(flow, t["_dense2.weight"], t["_dense2.bias"]))
        return F.linear(flow, t["_policy_head.weight"], t["_policy_head.bias"]), torch.tanh(F.linear(flow, t["_value_head.weight"], t["_value_head.bias"]))

    def bn(f_, prefix, affine):
        return F.batch_norm(f_, t[f"{prefix}.running_mean"], t[f"{prefix}.running_var"],
                            t[f"{prefix}.weight"] if affine else None, t[f"{prefix}.bias"] if affine else None,
                            training=False, eps=BN_EPS)

    flow = F.relu(bn(F.conv2d(flow, t["_conv1._conv.weight"], padding=1), "_conv1._bn", True))
    for i in range(cfg.blocks):
        p = f"_residual_blocks.{i}"
        inp = flow
        flow = F.relu(bn(F.conv2d(inp, t[f"{p}._conv1.weight"], padding=1), f"{p}._bn1", False))
        flow = bn(F.conv2d(flow, t[f"{p}._conv2.weight"], padding=1), f"{p}._bn2", True)
        flow = F.relu(inp + flow)
    v = F.relu(bn(F.conv2d(flow, t["_value_head.0._conv.weight"]), "_value_head.0._bn", False))
    v = F.relu(F.linear(v.flatten(1), t["_value_head.2.weight"], t["_value_head.2.bias"]))
    v = torch.tanh(F.linear(v, t["_value_head.4.weight"], t["_value_head.4.bias"]))
    p_ = F.relu(bn(F.conv2d(flow, t["_policy_head.0._conv.weight"]), "_policy_head.0._bn", False))
    p_ = F.linear(p_.flatten(1), t["_policy_head.2.weight"], t["_policy_head.2.bias"])
    return p_, v
