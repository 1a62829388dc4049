"""Synthetic highD/NGSIM-shaped scenes.  One scene = one batch row of the reference's
collate output (reference scripts/train.py:301-347); normalisation follows train.py:164-184."""
import torch

LANE_SIZES = (33, 32, 22, 14)   # real lane-polygon point counts (reference scripts/graph.py:7-215)


def make_scenes(B, seq_len=15, out_len=25, vision_dim=512, max_points=64, l_text=128, vocab=32000, seed=1234,
                ragged_text=True, poly_sizes=LANE_SIZES):
    g = torch.Generator().manual_seed(seed)
    T = seq_len + out_len
    rx = 100.0 + 1400.0 * torch.rand(B, generator=g)          # x-range in px (windows <100 px are skipped, 172)
    ry = 5.0 + 75.0 * torch.rand(B, generator=g)
    x0 = 3840.0 * torch.rand(B, generator=g) * 0.5
    y0 = 700.0 + 600.0 * torch.rand(B, generator=g)
    t = torch.linspace(0, 1, T)[None, :]
    px = x0[:, None] + rx[:, None] * t + 2.0 * torch.randn(B, T, generator=g)
    py = y0[:, None] + ry[:, None] * (t ** 2) * torch.sign(torch.randn(B, 1, generator=g)) + 0.5 * torch.randn(B, T, generator=g)
    min_x, max_x = px.min(1).values, px.max(1).values
    min_y, max_y = py.min(1).values, py.max(1).values
    nx = (px - min_x[:, None]) / (max_x - min_x)[:, None]
    ny = (py - min_y[:, None]) / (max_y - min_y)[:, None]
    traj = torch.stack([nx, ny], dim=1)                        # (B,2,T)
    vision = torch.randn(B, seq_len, vision_dim, generator=g)
    vision = vision / vision.norm(dim=-1, keepdim=True)
    sizes = torch.tensor(poly_sizes)[torch.randint(0, len(poly_sizes), (B,), generator=g)]
    poly = torch.zeros(B, max_points, 2)
    pts = torch.rand(B, max_points, 2, generator=g)
    pts[..., 0] = pts[..., 0] * 3839.0
    pts[..., 1] = 700.0 + pts[..., 1] * 750.0
    keep = torch.arange(max_points)[None, :] < sizes[:, None]
    poly[keep] = pts[keep]
    ids = torch.randint(0, vocab, (B, l_text), generator=g)
    if ragged_text:
        lo = max(1, (3 * l_text) // 4)
        valid = torch.randint(lo, l_text + 1, (B,), generator=g)
    else:
        valid = torch.full((B,), l_text)
    mask = (torch.arange(l_text)[None, :] < valid[:, None]).long()
    return {
        "x": traj[:, :, :seq_len].contiguous().float(), "y": traj[:, :, seq_len:].contiguous().float(),
        "vision": vision.float(), "polygon": poly.float(), "poly_len": [int(s) for s in sizes],
        "norm_stat": [(float(a), float(b), float(c), float(d)) for a, b, c, d in zip(min_x, max_x, min_y, max_y)],
        "input_ids": ids, "attention_mask": mask,
        "context_str": ["synthetic scene"] * B,
    }
