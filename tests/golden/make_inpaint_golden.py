"""Generate tests/golden/inpaint.npz by running cv2 / Pillow exactly as A1111's processing.py calls them for the
request the reference sends (signerf/diffuser/diffuser.py:164-169: mask_blur 4, inpainting_fill 1, inpaint_full_res 0).
The A1111 lines being reproduced (modules/processing.py, StableDiffusionProcessingImg2Img.init / apply_overlay):

    np_mask = cv2.GaussianBlur(np.array(image_mask), (kernel_size, 1), self.mask_blur_x)   # kernel_size = 2*int(2.5*blur+0.5)+1
    np_mask = cv2.GaussianBlur(np_mask, (1, kernel_size), self.mask_blur_y)
    np_mask = np.clip(np_mask.astype(np.float32) * 2, 0, 255).astype(np.uint8)             # mask_for_overlay
    latmask = latent_mask.convert('RGB').resize((w // 8, h // 8)); latmask = np.around(np.array(latmask, float32)[..,0] / 255)
    image_masked = Image.new('RGBa', size); image_masked.paste(image.convert("RGBA").convert("RGBa"),
                                                               mask=ImageOps.invert(self.mask_for_overlay.convert('L')))
    overlay = image_masked.convert('RGBA');  image = image.convert('RGBA'); image.alpha_composite(overlay); image.convert('RGB')

Rerun with `python tests/golden/make_inpaint_golden.py` (needs cv2 + Pillow; versions are stored in the fixture)."""
import os

import cv2
import numpy as np
import PIL
from PIL import Image, ImageOps

OUT = os.path.dirname(os.path.abspath(__file__))


def blobs(rng, h, w, n):
    m = np.zeros((h, w), np.uint8)
    for _ in range(n):
        cy, cx, ry, rx = rng.integers(0, h), rng.integers(0, w), rng.integers(3, h // 3), rng.integers(3, w // 3)
        yy, xx = np.ogrid[:h, :w]
        m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = 255
    return m


def main():
    rng = np.random.default_rng(7)
    out = {"cv2_version": cv2.__version__, "pil_version": PIL.__version__}
    cases = [(64, 96, 3), (40, 40, 2), (128, 256, 5), (17, 23, 1)]
    for ci, (h, w, n) in enumerate(cases):
        mask = blobs(rng, h, w, n)
        if ci == 3:
            mask = rng.integers(0, 256, (h, w), dtype=np.uint8)      # arbitrary grey levels, odd size
        for blur in (4, 2):
            ks = 2 * int(2.5 * blur + 0.5) + 1
            bx = cv2.GaussianBlur(mask, (ks, 1), blur)
            bxy = cv2.GaussianBlur(bx, (1, ks), blur)
            out[f"c{ci}_blur{blur}_x"] = bx
            out[f"c{ci}_blur{blur}_xy"] = bxy
        blurred = out[f"c{ci}_blur4_xy"]
        overlay_mask = np.clip(blurred.astype(np.float32) * 2, 0, 255).astype(np.uint8)
        out[f"c{ci}_mask"] = mask
        out[f"c{ci}_overlay_mask"] = overlay_mask
        lh, lw = max(1, h // 8), max(1, w // 8)
        lat_u8 = np.array(Image.fromarray(blurred).convert("RGB").resize((lw, lh)))[..., 0]
        out[f"c{ci}_lat_u8"] = lat_u8
        out[f"c{ci}_latmask"] = np.around(lat_u8.astype(np.float32) / 255).astype(np.float32)
        orig = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        gen = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        image = Image.fromarray(orig)
        image_masked = Image.new("RGBa", (w, h))
        image_masked.paste(image.convert("RGBA").convert("RGBa"), mask=ImageOps.invert(Image.fromarray(overlay_mask).convert("L")))
        overlay = image_masked.convert("RGBA")
        res = Image.fromarray(gen).convert("RGBA")
        res.alpha_composite(overlay)
        out[f"c{ci}_orig"], out[f"c{ci}_gen"] = orig, gen
        out[f"c{ci}_composited"] = np.array(res.convert("RGB"))
    # generic PIL bicubic resizes (other ratios, both axes)
    for ri, (h, w, oh, ow) in enumerate([(64, 64, 8, 8), (100, 60, 13, 7), (33, 47, 33, 6), (24, 24, 48, 30)]):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"r{ri}_in"] = img
        out[f"r{ri}_out"] = np.array(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
    # Gaussian taps recovered from cv2 itself: response to an impulse of every grey level pins each Q8 tap
    for sigma, ks in ((4.0, 21), (2.0, 11), (1.5, 9)):
        row = np.zeros((1, 4 * ks + 1), np.uint8)
        row[0, 2 * ks] = 255
        out[f"impulse_s{sigma}_k{ks}"] = cv2.GaussianBlur(row, (ks, 1), sigma)
    np.savez_compressed(os.path.join(OUT, "inpaint.npz"), **out)
    print("wrote inpaint.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
