"""TEST INFRASTRUCTURE: mints tests/golden/collate_b5.pt by running the UNMODIFIED reference `custom_collate_fn`
(scripts/train.py:301-347, imported through oracle/ref_loader.py) on seeded synthetic dataset samples.
    python -m oracle.make_collate_golden"""
import os

import torch

from .ref_loader import load_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def samples(n=5, t_in=15, t_out=25, seed=11):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        lt = [40, 33, 57, 12, 57][i % 5]
        plen = [33, 32, 22, 14, 64][i % 5]
        poly = torch.zeros(64, 2)
        poly[:plen] = torch.rand(plen, 2, generator=g) * 1000
        out.append(dict(traj_emb=torch.rand(t_in, 2, generator=g), target_traj=torch.rand(t_out, 2, generator=g),
                        vision_emb=torch.randn(t_in, 512, generator=g), lane_polygon=poly, lane_polygon_len=plen,
                        norm_stat=(float(100 + i), float(900 + i), float(700 + i), float(760 + i)), context_str=f"ctx {i}", answer_str=f"ans {i}",
                        track_id=100 + i, input_ids=torch.randint(0, 32000, (lt,), generator=g),
                        attention_mask=torch.ones(lt, dtype=torch.int64), labels=torch.randint(0, 32000, (lt,), generator=g)))
    return out


def main():
    ref = load_reference("scripts/train.py", llama_cfg=None)
    s = samples()
    want = ref.custom_collate_fn(s)
    torch.save({"samples": s, "collated": want}, os.path.join(ROOT, "tests", "golden", "collate_b5.pt"))
    print({k: (tuple(v.shape), v.dtype) if torch.is_tensor(v) else type(v).__name__ for k, v in want.items()})


if __name__ == "__main__":
    main()
