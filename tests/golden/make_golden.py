"""Generate tests/golden/*.npz by running the REFERENCE's own code (the parts importable without
nerfstudio) and cv2 in the build container.  /root/reference does not exist on the GPU box, so the
outputs are committed as fixtures; rerun with `python tests/golden/make_golden.py` to regenerate.

Pinned here (reference file:line):
  signerf/utils/intersection.py:5-56        intersect_with_aabb
  signerf/utils/poses_generation.py:22-73   circle_poses (the benchmark camera ring)
  signerf/utils/image_tensor_converter.py:7-33  tensor_to_image (uint8 truncation)
  cv2.getStructuringElement / cv2.dilate    as called at signerf/datasetgenerator/datasetgenerator.py:775-778
  torch F.interpolate bilinear              as called at datasetgenerator.py:526-528, :585
"""
import os
import sys
import warnings

import cv2
import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
from signerf.utils.image_tensor_converter import tensor_to_image  # noqa: E402
from signerf.utils.intersection import intersect_with_aabb  # noqa: E402
from signerf.utils.poses_generation import circle_poses  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(1234)
    # --- aabb
    H, W = 24, 40
    o = (torch.rand(H, W, 3, generator=g) - 0.5) * 1.5
    d = torch.randn(H, W, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    aabb = torch.tensor([[-0.1, -0.15, -0.2], [0.1, 0.25, 0.3]])
    nears, fars = intersect_with_aabb(o, d, aabb)
    np.savez(os.path.join(OUT, "aabb.npz"), o=o.numpy(), d=d.numpy(), aabb=aabb.numpy(), nears=nears.numpy(),
             fars=fars.numpy())
    # --- poses (benchmark ring: GUI default, interface.py:62-71)
    poses16 = circle_poses(16, "cpu", 0.5, 90, (0, 300), [0, 0, 0], [0, 0, 0])
    poses5 = circle_poses(5, "cpu", 0.5, 60, (0, 300), [0.1, 0.0, -0.1], [0, 0, 0])
    np.savez(os.path.join(OUT, "poses.npz"), poses16=poses16.numpy(), poses5=poses5.numpy())
    # --- quantisation
    x = torch.cat([torch.rand(1000, generator=g), torch.tensor([0.0, 0.999, 1.0, 0.5, 254.5 / 255, 1.0 / 255, 1.001, 1.5])])
    x3 = x.reshape(-1, 1, 1).repeat(1, 1, 3)
    q = np.array(tensor_to_image(x3))
    np.savez(os.path.join(OUT, "quantize.npz"), x=x3.numpy(), q=q)
    # --- dilation
    se50 = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (50, 50))
    se_odd = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 11))
    se_small = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (4, 4))
    masks, outs50, outs711 = [], [], []
    for i, p in enumerate([0.0005, 0.002, 0.0]):
        m = (torch.rand(96, 128, generator=g) < p).numpy().astype(float)
        if i == 1:
            m[0, 0] = 1.0
            m[95, 127] = 1.0
            m[40:44, 60:70] = 1.0
        masks.append(m)
        outs50.append(cv2.dilate(m, se50) > 0)
        outs711.append(cv2.dilate(m, se_odd) > 0)
    np.savez_compressed(os.path.join(OUT, "dilate.npz"), se50=se50, se_7x11=se_odd, se_4x4=se_small,
                        masks=np.stack(masks).astype(np.uint8), out50=np.stack(outs50), out7x11=np.stack(outs711))
    # --- bilinear resize (as F.interpolate is called on permuted HWC tensors)
    img = torch.rand(20, 28, 3, generator=g)
    def interp(t, h, w):
        return F.interpolate(t.permute(2, 0, 1).unsqueeze(0), (h, w), mode="bilinear", align_corners=False).squeeze(0).permute(1, 2, 0)
    np.savez(os.path.join(OUT, "resize.npz"), img=img.numpy(), down=interp(img, 10, 14).numpy(),
             odd=interp(img, 13, 9).numpy(), up=interp(img, 40, 56).numpy(), same=interp(img, 20, 28).numpy())
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
