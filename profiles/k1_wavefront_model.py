"""Per-level L1 cost of K1's hash-grid gathers, from the address pattern itself (no GPU needed).

K1 marches 32 rays (an 8 x 4 pixel tile) per warp in lock step; every level issues 8 LDG.64 (one per cell corner).  An
LDG costs one L1 tag request / data wavefront per distinct 128-B line its 32 lanes touch, and moves one 32-B sector per
distinct sector.  This script replays the benchmark's geometry (camera ring, 512 x 512, flat 128 piecewise bins) through
the oracle's hash function for a sample of warps and prints, per level, the mean distinct lines and sectors per gather
instruction - the per-level breakdown ncu cannot give (levels share SASS instructions: the level loop is rolled 4 x 4).
    python profiles/k1_wavefront_model.py > profiles/r2_k1_wavefronts_per_level.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerfacto_ref as R  # noqa: E402
from tests.helpers import ring_cameras  # noqa: E402

torch.manual_seed(0)
H = W = 512
S = 128
m = R.make_model(0)
enc = m.field.encoding
c2w, intr = ring_cameras(16, W, H)
bins = R.flat_bin_edges(S, m.near, m.far)
mids = (bins[:-1] + bins[1:]) / 2
rng = np.random.default_rng(0)
L, T = 16, enc.hash_table_size
lines = np.zeros(L)
sectors = np.zeros(L)
coinc = np.zeros(L)
n_inst = 0
for _ in range(48):                                   # 48 random warps x all 128 samples
    v = int(rng.integers(0, 16))
    rays = R.generate_rays(c2w[v], *intr[v].tolist(), W, H)
    tx, ty = int(rng.integers(0, W // 8)), int(rng.integers(0, H // 4))
    ids = torch.tensor([(ty * 4 + r // 8) * W + tx * 8 + r % 8 for r in range(32)])
    o, d = rays.origins[ids], rays.directions[ids]
    pos = o[:, None, :] + d[:, None, :] * mids[None, :, None]          # [32, S, 3]
    p, _ = m.field.contracted_positions(pos)
    idx, _ = enc.corner_indices(p.reshape(-1, 3))                        # [32*S, L, 8] absolute table rows
    idx = idx.view(32, S, L, 8).numpy()
    addr = idx.astype(np.int64) * 8
    for lvl in range(L):
        a = addr[:, :, lvl, :]                                          # [32 lanes, S, 8 corners]
        ln = np.sort(a // 128, axis=0)
        sc = np.sort(a // 32, axis=0)
        lines[lvl] += (1 + (np.diff(ln, axis=0) != 0).sum(axis=0)).sum()
        sectors[lvl] += (1 + (np.diff(sc, axis=0) != 0).sum(axis=0)).sum()
    n_inst += S * 8
res = enc.scalings.tolist()
print("# K1 hash-grid gathers: distinct 128-B lines (= L1 tag requests / data wavefronts) and 32-B sectors per LDG.64 of a warp")
print("# benchmark geometry: 16-view ring, 512 x 512, flat 128 bins, warp = 8 x 4 pixel tile; 48 random warps x 128 samples")
print(f"{'level':>5} {'res':>6} {'lines/LDG':>10} {'sectors/LDG':>12} {'share of lines':>15}")
tot = lines.sum()
for lvl in range(L):
    print(f"{lvl:5d} {int(res[lvl]):6d} {lines[lvl] / n_inst:10.2f} {sectors[lvl] / n_inst:12.2f} {100 * lines[lvl] / tot:14.1f}%")
print(f"  all        {tot / n_inst / L:10.2f} {sectors.sum() / n_inst / L:12.2f}   (mean per LDG over the 16 levels)")
print(f"# per sample (32 lanes): {tot / n_inst * 8:.0f} tag requests for {L * 8} LDG; the ideal for a fully coherent warp is {L * 8}")
print(f"# levels 0-4 (the 'dense re-indexable' ones of SURVEY section 7) carry {100 * lines[:5].sum() / tot:.1f}% of the L1 work;"
      f" levels 8-15 carry {100 * lines[8:].sum() / tot:.1f}%")
