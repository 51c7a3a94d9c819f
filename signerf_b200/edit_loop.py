"""BASELINE config 5 — "full SIGNeRF edit loop: proxy mesh, 30 cameras, refinement rounds": dataset generation
(plugin.DatasetGenerator.generate_dataset, reference datasetgenerator.py:185-393) alternating with NeRF fine-tuning on the
generated images (the reference's trainer swaps the pipeline onto the generated dataset and trains SIGNeRFModel on it,
signerf_trainer.py:219-235; here the training step of signerf_b200/train.py: with a `NerfactoTrainer` the whole nerfacto
step - proposal sampler in training mode, rgb + interlevel + distortion losses, per-image appearance embeddings, Adam on
fields and proposal networks; with a `FieldTrainer` the main-field slice on flat bins).  Multi-GPU: generation shards the
dataset cameras over the ranks; training is data-parallel - every rank draws its own patches and the gradients are summed
with ONE flattened all-reduce per step (the path's only reduction)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import ops
from .train import (FieldTrainer, NerfactoTrainer, PatchPixelSampler, PatchPixelSamplerConfig, ProposalUpdateSchedule,
                    proposal_anneal)


def load_generated_images(dataset_dir) -> Tuple[Tensor, Tensor, Tensor]:
    """The generated dataset as the reference's dataparser reads it back (signerf_dataparser.py:99-146): images [N,H,W,3]
    fp32 0..1 (`image_to_tensor`: uint8 / 255), scene-space c2w [N,3,4], intrinsics [N,4], from transforms.json."""
    from PIL import Image
    root = Path(dataset_dir)
    meta = json.loads((root / "transforms.json").read_text())
    imgs, c2w, intr = [], [], []
    for fr in meta["frames"]:
        im = Image.open(root / fr["file_path"])
        if im.mode == "RGBA":
            im = im.convert("RGB")
        imgs.append(torch.from_numpy(np.array(im, dtype="float32") / 255.0))
        c2w.append(torch.tensor(fr["scene_transform_matrix"] if "scene_transform_matrix" in fr else fr["transform_matrix"],
                                dtype=torch.float32)[:3, :4])
        intr.append(torch.tensor([fr["fl_x"], fr["fl_y"], fr["cx"], fr["cy"]], dtype=torch.float32))
    return torch.stack(imgs), torch.stack(c2w), torch.stack(intr)


class FineTuner:
    """K fine-tune steps on a set of posed images."""

    def __init__(self, trainer: FieldTrainer, num_samples: int = 48, rays_per_batch: int = 16384, patch_size: int = 32,
                 near: float = 0.05, far: float = 1000.0, seed: int = 0):
        self.trainer, self.S = trainer, num_samples
        dev = trainer.field.device
        self.sampler = PatchPixelSampler(PatchPixelSamplerConfig(patch_size=patch_size, num_rays_per_batch=rays_per_batch))
        self.bins = ops.piecewise_bin_edges(num_samples, near, far).to(dev)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        # nerfacto's step-dependent callbacks; the reference's trainer resets the step count when it switches to the generated
        # dataset (signerf_trainer.py:321-325), so every fit() call of a fresh FineTuner starts them at step 0
        self.step, self.schedule = 0, ProposalUpdateSchedule()

    def fit(self, images: Tensor, c2w: Tensor, intr: Tensor, steps: int) -> List[float]:
        """images [N,H,W,3] / cameras on the field's device.  Returns the loss of every `max(1, steps // 10)`-th step."""
        import torch.distributed as dist
        tr = self.trainer
        dev = tr.field.device
        images = images.to(dev)
        n, h, w = images.shape[0], images.shape[1], images.shape[2]
        origins, dirs, _, _ = ops.generate_rays(c2w.to(dev), intr.to(dev), h, w)          # [N,H,W,3] each
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        losses = []
        full = isinstance(tr, NerfactoTrainer)
        for it in range(steps):
            idx = self.sampler.sample_method(self.sampler.num_rays_per_batch, n, h, w, device=dev, generator=self.gen)
            i, y, x = idx[:, 0], idx[:, 1], idx[:, 2]
            o, d, target = origins[i, y, x].contiguous(), dirs[i, y, x].contiguous(), images[i, y, x].contiguous()
            if full:
                jitter = torch.rand((3, o.shape[0]), device=dev, generator=self.gen)
                cams = i.to(torch.int32) if tr.embedding is not None and tr.embedding.shape[0] >= n else None
                out = tr.forward_backward(o, d, target, jitter, cams, anneal=proposal_anneal(self.step),
                                          update_proposals=self.schedule(self.step))
                self.step += 1
                if world > 1:
                    grads = tr.all_gradients()
                    flat = torch.cat([g.reshape(-1) for g in grads])
                    dist.all_reduce(flat)
                    flat.mul_(1.0 / world)
                    off = 0
                    for g in grads:
                        g.copy_(flat[off:off + g.numel()].view_as(g))
                        off += g.numel()
                tr.optimizer_step()
                loss = sum(out.values())
            elif world == 1:
                loss = tr.step(o, d, self.bins, target)
            else:
                from . import train as T
                tr.zero_grad()
                rgb, _, saved = T.train_forward(tr.field, o, d, self.bins)
                loss, grad = T.rgb_loss(rgb, target, tr.use_l1)
                tr.backward(o, d, self.bins, saved, grad)
                for g in (tr.grad_table, tr.grad_mlp):           # data-parallel: mean of the ranks' mean-reduced losses
                    dist.all_reduce(g)
                    g.mul_(1.0 / world)
                tr.optimizer_step()
            if it % max(1, steps // 10) == 0 or it == steps - 1:
                losses.append(float(loss))
        tr.refresh_renderer()
        return losses
