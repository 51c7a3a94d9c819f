"""Host-side handle of a nerfacto field living on the GPU.

Holds the torch tensors (hash tables stay where nerfstudio put them — the library keeps a
reference, it never copies 64 MiB tables) and the opaque `SgnField*` created from them.  This is the
object `SIGNeRFModel` hands to the fused renderer in place of running NerfactoModel.get_outputs
(reference signerf/signerf.py:27-39, call site signerf/datasetgenerator/datasetgenerator.py:694).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@dataclass
class HashGridParams:
    table: Tensor            # [L * 2^log2, 2] fp32 CUDA
    scalings: Tensor         # [L] fp32 (HashEncoding.scalings, passed through untouched)
    log2_size: int

    @property
    def num_levels(self) -> int:
        return int(self.scalings.numel())


@dataclass
class LinearParams:
    weight: Tensor  # [out, in]
    bias: Tensor    # [out]


class NerfactoFieldB200:
    """Parameters of NerfactoField (+ optional 2 proposal networks) bound to a `SgnField` handle."""

    def __init__(self, grid: HashGridParams, base: Sequence[LinearParams], head: Sequence[LinearParams],
                 appearance_mean: Tensor, average_init_density: float,
                 prop_grids: Sequence[HashGridParams] = (), prop_mlps: Sequence[Sequence[LinearParams]] = ()):
        if not grid.table.is_cuda:
            raise RuntimeError("signerf_b200 needs the hash tables on a CUDA device (no CPU fallback)")
        self.device = grid.table.device
        self.grid = grid
        self.base, self.head = list(base), list(head)
        self.prop_grids, self.prop_mlps = list(prop_grids), [list(m) for m in prop_mlps]
        self.appearance_mean = appearance_mean
        self.average_init_density = float(average_init_density)
        self._keep: List[object] = []  # host arrays / tensors that must outlive the create call
        self._handle = C.c_void_p()
        self._create()

    # -- construction helpers -------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, sd: Mapping[str, Tensor], scalings: Tensor, log2_size: int = 19,
                        average_init_density: float = 0.01, prop_scalings: Sequence[Tensor] = (),
                        prop_log2: Sequence[int] = (17, 17), device: str = "cuda") -> "NerfactoFieldB200":
        """Build from a nerfacto (`implementation="torch"`) pipeline checkpoint's `_model.` tensors.

        Key names follow nerfstudio 1.0.x (to be re-verified against a real checkpoint):
        field.mlp_base_grid.hash_table, field.mlp_base_mlp.layers.{0,1}.*, field.mlp_head.layers.{0,1,2}.*,
        field.embedding_appearance.embedding.weight, proposal_networks.{i}.encoding.hash_table,
        proposal_networks.{i}.mlp_base.layers.*.  SIGNeRF drops the appearance table on load
        (signerf/signerf_pipeline.py:110-111) so eval uses the mean of whatever table the model holds.
        """
        def lin(prefix: str) -> LinearParams:
            return LinearParams(sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float())

        def first(*names):
            for n in names:
                if n in sd:
                    return n
            raise KeyError(f"none of {names} in state dict")

        tab = sd[first("field.mlp_base_grid.hash_table", "field.mlp_base.encoder.hash_table",
                       "field.encoding.hash_table")]
        basep = "field.mlp_base_mlp.layers" if "field.mlp_base_mlp.layers.0.weight" in sd else "field.mlp_base.layers"
        grid = HashGridParams(tab.to(device=device, dtype=torch.float32).contiguous(), scalings.float().cpu(), log2_size)
        base = [lin(f"{basep}.{i}") for i in range(2)]
        head = [lin(f"field.mlp_head.layers.{i}") for i in range(3)]
        app = sd["field.embedding_appearance.embedding.weight"].float().mean(dim=0)
        pg, pm = [], []
        for i, sc in enumerate(prop_scalings):
            t = sd[first(f"proposal_networks.{i}.encoding.hash_table", f"proposal_networks.{i}.mlp_base.encoder.hash_table")]
            pg.append(HashGridParams(t.to(device=device, dtype=torch.float32).contiguous(), sc.float().cpu(), prop_log2[i]))
            pp = first(f"proposal_networks.{i}.mlp.layers.0.weight", f"proposal_networks.{i}.mlp_base.layers.0.weight",
                       f"proposal_networks.{i}.mlp_base_mlp.layers.0.weight").rsplit(".0.weight", 1)[0]
            pm.append([lin(f"{pp}.{j}") for j in range(2)])
        return cls(grid, base, head, app, average_init_density, pg, pm)

    # -- C ABI ----------------------------------------------------------------------------
    def _host(self, t: Tensor) -> np.ndarray:
        a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
        self._keep.append(a)
        return a

    def _grid_desc(self, g: HashGridParams) -> _lib.SgnHashGrid:
        if g.table.dtype != torch.float32 or not g.table.is_contiguous():
            raise ValueError("hash table must be contiguous float32")
        rows = g.num_levels * (1 << g.log2_size)
        if tuple(g.table.shape) != (rows, 2):
            raise ValueError(f"hash table shape {tuple(g.table.shape)} != ({rows}, 2)")
        sc = self._host(g.scalings)
        return _lib.SgnHashGrid(C.c_void_p(g.table.data_ptr()), _fptr(sc), g.num_levels, g.log2_size)

    def _lin_desc(self, l: LinearParams) -> _lib.SgnLinear:
        w, b = self._host(l.weight), self._host(l.bias)
        return _lib.SgnLinear(_fptr(w), _fptr(b), int(w.shape[1]), int(w.shape[0]))

    def _create(self) -> None:
        lib = _lib.load()
        d = _lib.SgnFieldDesc()
        d.grid = self._grid_desc(self.grid)
        for i in range(2):
            d.base[i] = self._lin_desc(self.base[i])
        for i in range(3):
            d.head[i] = self._lin_desc(self.head[i])
        d.h_appearance = _fptr(self._host(self.appearance_mean))
        d.average_init_density = self.average_init_density
        d.num_proposals = len(self.prop_grids)
        for i, g in enumerate(self.prop_grids):
            d.prop_grid[i] = self._grid_desc(g)
            for j in range(2):
                d.prop_mlp[i][j] = self._lin_desc(self.prop_mlps[i][j])
        with torch.cuda.device(self.device):
            _lib.check(lib.sgn_field_create(C.byref(d), C.byref(self._handle)))

    @property
    def handle(self) -> C.c_void_p:
        if not self._handle:
            raise RuntimeError("field handle already destroyed")
        return self._handle

    def close(self) -> None:
        if self._handle:
            _lib.load().sgn_field_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
