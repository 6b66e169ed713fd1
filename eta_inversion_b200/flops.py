"""Analytic work of one SD-1.x UNet forward per batch row (dense contractions only), by kernel category.
Mirrors the layer walk of csrc/unet.cu; used for roofline arithmetic (401.637 GMAC/row)."""
from __future__ import annotations

from typing import Dict, Sequence


def unet_macs_per_row(channels: Sequence[int] = (320, 640, 1280, 1280), hw: int = 64, ctx_len: int = 77,
                      cross_dim: int = 768, heads: int = 8) -> Dict[str, float]:
    c = list(channels)
    temb = 4 * c[0]
    m = {"conv3x3": 0.0, "gemm": 0.0, "self_attn": 0.0, "cross_attn": 0.0}

    def conv(n, cin, cout):
        m["conv3x3"] += n * 9 * cin * cout

    def lin(n, k, o):
        m["gemm"] += n * k * o

    def res(n, cin, cout):
        conv(n, cin, cout)
        conv(n, cout, cout)
        lin(1, temb, cout)
        if cin != cout:
            lin(n, cin, cout)

    def tfm(n, C):
        lin(n, C, C)            # proj_in
        lin(n, C, 3 * C)        # qkv
        m["self_attn"] += 2 * n * n * C
        lin(n, C, C)            # to_out
        lin(n, C, C)            # cross q
        lin(ctx_len, cross_dim, 2 * C)
        m["cross_attn"] += 2 * n * ctx_len * C
        lin(n, C, C)            # cross out
        lin(n, C, 8 * C)        # GEGLU
        lin(n, 4 * C, C)
        lin(n, C, C)            # proj_out

    n = hw * hw
    conv(n, 4, c[0])
    lin(1, c[0], temb)
    lin(1, temb, temb)
    cout = c[0]
    for i in range(4):
        cin, cout = cout, c[i]
        for j in range(2):
            res(n, cin if j == 0 else cout, cout)
            if i < 3:
                tfm(n, cout)
        if i < 3:
            n //= 4
            conv(n, cout, cout)
    res(n, c[3], c[3]); tfm(n, c[3]); res(n, c[3], c[3])
    rev = c[::-1]
    cout = rev[0]
    for i in range(4):
        prev, cout = cout, rev[i]
        cin = rev[min(i + 1, 3)]
        for j in range(3):
            res(n, (prev if j == 0 else cout) + (cin if j == 2 else cout), cout)
            if i > 0:
                tfm(n, cout)
        if i < 3:
            n *= 4
            conv(n, cout, cout)
    conv(n, c[0], 4)
    m["total"] = sum(m.values())
    return m


if __name__ == "__main__":
    b = unet_macs_per_row()
    print({k: round(v / 1e9, 3) for k, v in b.items()})
