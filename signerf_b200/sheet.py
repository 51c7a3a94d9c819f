"""Reference-sheet assembly on the GPU: the body of reference
`DatasetGenerator.generate_reference_sheet` up to the diffuser call
(signerf/datasetgenerator/datasetgenerator.py:498-539) with every per-view loop batched into one launch per
kernel: K1 render -> K2/K3 mask+condition -> K4 paste.  No host synchronisation inside."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops
from .field import NerfactoFieldB200


@dataclass
class SheetBuffers:
    image: Tensor      # [Hs, Ws, 3] fp32, background 1
    mask: Tensor       # [Hs, Ws, 1] fp32 0/1
    condition: Tensor  # [Hs, Ws, 1] fp32


class ReferenceSheetRenderer:
    """Renders V views and assembles image / mask / condition sheets.

    `views_rgb/depth/mask/cond` of the last call stay available for the caller (references[i]["render"] etc.).
    """

    def __init__(self, fld: NerfactoFieldB200, layout: ops.SheetLayout, height: int, width: int,
                 render_opts: ops.RenderOptions, mask_opts: ops.MaskOptions):
        self.field, self.layout = fld, layout
        self.height, self.width = height, width
        self.render_opts, self.mask_opts = render_opts, mask_opts
        dev = fld.device
        self.buffers = SheetBuffers(
            torch.ones((layout.height, layout.width, 3), dtype=torch.float32, device=dev),
            torch.zeros((layout.height, layout.width, 1), dtype=torch.float32, device=dev),
            torch.zeros((layout.height, layout.width, 1), dtype=torch.float32, device=dev))
        self.views: Optional[Tuple[Tensor, Tensor, Tensor, Tensor]] = None

    def render_tiles(self, c2w: Tensor, intr: Tensor):
        """K1 + K2/K3 for V views: (rgb [V,H,W,3], depth, mask uint8, cond)."""
        rgb, depth = ops.render_views(self.field, c2w, intr, self.height, self.width, self.render_opts)
        mask, cond, _ = ops.mask_condition(c2w, intr, depth, self.mask_opts)
        self.views = (rgb, depth, mask, cond)
        return self.views

    def paste(self, rgb: Tensor, mask: Tensor, cond: Tensor, first_cell: int = 0) -> SheetBuffers:
        """K4: resize + paste tiles starting at grid cell `first_cell` (row-major)."""
        b = self.buffers
        ops.sheet_paste(rgb, b.image, self.layout, first_cell)
        ops.sheet_paste(mask, b.mask, self.layout, first_cell, threshold=0.5)
        ops.sheet_paste(cond, b.condition, self.layout, first_cell)
        return b

    def __call__(self, c2w: Tensor, intr: Tensor, first_cell: int = 0) -> SheetBuffers:
        rgb, _, mask, cond = self.render_tiles(c2w, intr)
        return self.paste(rgb, mask, cond, first_cell)
