"""CPU, gloo ranks: the multi-GPU PRODUCT path — `plugin.DatasetGenerator.generate_dataset` under torch.distributed.

Hot loop #1 (reference datasetgenerator.py:331-338) is sharded camera i -> rank i mod world, the reference views
per view with one all-gather, the single reference-sheet diffusion runs on rank 0 and is broadcast, rank 0 writes
transforms.json.  The CUDA kernels are replaced by the ORACLE's torch restatements (oracle/sheet_ref.py) and a
deterministic stand-in renderer, so what is tested here is the partition / gather / merge logic: the dataset written
by 2 and 3 ranks must be byte-identical to the one a single process writes."""
import hashlib
import os
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

H, W, ROWS, COLS, DS = 16, 24, 2, 2, 2
N_GEN = 5


def _patch_kernels():
    """CPU stand-ins for the sheet kernels (test infrastructure: the oracle's restatement of the same reference lines)."""
    import torch.nn.functional as F
    from oracle import sheet_ref as S
    from signerf_b200 import ops

    def sheet_paste(src, sheet, layout, first_cell=0, threshold=None):
        for v in range(src.shape[0]):
            cell = first_cell + v
            r0 = (cell // layout.cols) * (layout.tile_h + layout.border)
            c0 = (cell % layout.cols) * (layout.tile_w + layout.border)
            t = S._interp(src[v].float(), layout.tile_h, layout.tile_w)
            sheet[r0:r0 + layout.tile_h, c0:c0 + layout.tile_w] = (t > threshold).float() if threshold is not None else t

    ops.sheet_paste = sheet_paste
    ops.sheet_cut = lambda sheet, layout, cell, h, w: S.cut_tile(sheet, cell, layout.cols, layout.tile_h, layout.tile_w, layout.border, h, w)
    ops.blend_masked = lambda e, b, m: S.blend(e, b, m)


def _make_generator(root: Path):
    import signerf_b200.plugin as P

    class Gen(P.DatasetGenerator):
        def _fused_graph(self, graph):
            return graph

        def render_views(self, graph, cameras, combine_shape_with_depth=None):     # deterministic function of the pose
            cam = P.base.as_camera_batch(cameras)
            out = []
            for c2w in cam.camera_to_worlds:
                g = torch.Generator().manual_seed(int(round(float(c2w[0, 3]) * 1000)))
                out.append((torch.rand(H, W, 3, generator=g), torch.rand(H, W, 1, generator=g) > 0.5, torch.rand(H, W, 1, generator=g)))
            return torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out]), torch.stack([o[2] for o in out])

    class Diff(P.Diffuser):
        calls = 0

        def diffuse(self, original_image, rendered_image, mask_image=None, condition_image=None):
            Diff.calls += 1
            return 1.0 - 0.5 * original_image

    cfg = P.DatasetGeneratorConfig(path=root, dataset_name="exp", rows=ROWS, cols=COLS, width=W, height=H, downscale_factor=DS,
                                   fx=float(W), fy=float(W), cx=W / 2, cy=H / 2)
    gen = Gen(cfg, torch.eye(4)[:3], 1.0, lambda x: x, "cpu")
    gen.diffuser = Diff(cfg.diffuser, "cpu")
    return gen, Diff


def _poses(n, offset):
    c2w = torch.eye(4)[:3].repeat(n, 1, 1)
    c2w[:, 0, 3] = torch.arange(n) * 0.001 + offset
    return c2w


def _run(root: Path):
    _patch_kernels()
    gen, Diff = _make_generator(root)
    graph = type("G", (), {"device": torch.device("cpu"), "render_aabb": None, "eval": lambda s: s, "train": lambda s: s})()
    gen.generate_dataset(graph, _poses(ROWS * COLS - 1, 0.1), synthetic_camera_to_worlds=_poses(N_GEN, 0.5))
    return Diff.calls


def _digest(root: Path):
    out = {}
    for p in sorted((root / "exp").rglob("*")):
        if p.is_file() and p.name != "config.yml":
            out[str(p.relative_to(root))] = hashlib.sha256(p.read_bytes()).hexdigest()
    return out


def _worker(rank: int, world: int, port: int, root: str, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = _run(Path(root))
    q.put((rank, calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_generate_dataset_writes_the_single_process_dataset(tmp_path, world):
    single = tmp_path / "single"
    assert _run(single) == 1 + N_GEN                          # reference sheet + one sheet per dataset camera
    ref = _digest(single)
    assert len([k for k in ref if "/images/" in k]) == ROWS * COLS - 1 + N_GEN
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    root = tmp_path / f"w{world}"
    port = 29700 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(root), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    # the reference sheet is diffused ONCE (rank 0) and broadcast; the dataset cameras split i mod world
    assert res[0] == 1 + len(range(0, N_GEN, world))
    for r in range(1, world):
        assert res[r] == len(range(r, N_GEN, world))
    assert _digest(root) == ref                               # every PNG and transforms.json byte-identical to 1 process
