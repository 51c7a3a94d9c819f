"""the generated-dataset directory / transforms.json writer and reader
(signerf/datasetgenerator/datasetgenerator.py:146-182, :286-295, :398-468; signerf/data/signerf_dataparser.py:99-146).
CPU: schema, key order, PNG pixels = tensor_to_image's truncation (golden from the reference function), read-back.
GPU: plugin.DatasetGenerator.generate_dataset end to end with a stand-in diffuser."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from signerf_b200.plugin import dataset_io as IO

FRAME_KEYS = ["fl_x", "fl_y", "cx", "cy", "w", "h", "file_path", "_mask_path", "transform_matrix", "scene_transform_matrix"]
TOP_KEYS = ["camera_model", "orientation_override", "method", "is_synthetic", "is_combined", "frames",
            "original_transform_matrix", "original_scale_factor"]


def test_png_pixels_match_tensor_to_image(golden_dir, tmp_path):
    q = np.load(os.path.join(golden_dir, "quantize.npz"))          # x [N,1,3] -> reference tensor_to_image pixels
    x = torch.tensor(q["x"])
    assert np.array_equal(IO.quantize_image(x), q["q"])
    w = IO.AsyncImageWriter(2)
    w.save(x, tmp_path / "rgb.png")
    w.save(x[..., :1], tmp_path / "grey.png")
    w.save((x[..., :1] > 0.5), tmp_path / "mask.png")
    w.close()
    assert np.array_equal(np.array(Image.open(tmp_path / "rgb.png")), q["q"])
    grey = Image.open(tmp_path / "grey.png")
    assert grey.mode == "L" and np.array_equal(np.array(grey), q["q"][..., 0])
    assert set(np.unique(np.array(Image.open(tmp_path / "mask.png")))) <= {0, 255}
    with pytest.raises(AssertionError):
        IO.quantize_image(torch.zeros(4, 4))


def test_writer_layout_schema_and_read_back(tmp_path):
    wr = IO.DatasetWriter(tmp_path, "exp", 2)
    wr.init_directory()
    for d in ("images", "masks", "conditions", "rendered", "originals", "images_2", "masks_2", "conditions_2", "rendered_2",
              "originals_2", "references"):
        assert (tmp_path / "exp" / d).is_dir()
    T = IO.DatasetWriter.new_transforms(True, False, torch.eye(4)[:3], 0.5)
    assert list(T) == TOP_KEYS
    g = torch.Generator().manual_seed(0)
    imgs = {"edited": torch.rand(8, 10, 3, generator=g), "render": torch.rand(8, 10, 3, generator=g),
            "mask": torch.rand(8, 10, 1, generator=g) > 0.5, "condition": torch.rand(8, 10, 1, generator=g),
            "edited_scaled": torch.rand(4, 5, 3, generator=g), "render_scaled": torch.rand(4, 5, 3, generator=g),
            "mask_scaled": torch.rand(4, 5, 1, generator=g) > 0.5, "condition_scaled": torch.rand(4, 5, 1, generator=g)}
    c2w = torch.tensor([[1.0, 0, 0, 0.1], [0, 1, 0, 0.2], [0, 0, 1, 0.3]])
    T = wr.save_generated_images(0, imgs, c2w, 10.0, 11.0, 5.0, 4.0, 10, 8, T)
    T = wr.save_generated_images(1, {"edited": imgs["edited"], "render": imgs["render"], "edited_scaled": imgs["edited_scaled"]},
                                 c2w, 10.0, 11.0, 5.0, 4.0, 10, 8, T, is_original=True)
    T["reference_indices"] = [0]
    T["generated_indices"] = [1]
    wr.write_transforms(T)
    wr.writer.close()
    text = (tmp_path / "exp" / "transforms.json").read_text()
    assert text == json.dumps(T, indent=4)                              # json.dump(transforms, file, indent=4)
    meta = json.loads(text)
    assert list(meta)[:8] == TOP_KEYS and list(meta["frames"][0]) == FRAME_KEYS
    f0 = meta["frames"][0]
    assert f0["file_path"] == "./images/image_0.png" and f0["_mask_path"] == "./masks/mask_0.png"
    assert f0["transform_matrix"] == f0["scene_transform_matrix"] and f0["scene_transform_matrix"][3] == [0.0, 0.0, 0.0, 1.0]
    assert (tmp_path / "exp" / "rendered" / "image_0.png").exists() and (tmp_path / "exp" / "originals" / "image_1.png").exists()
    assert np.array_equal(np.array(Image.open(tmp_path / "exp" / "images" / "image_0.png")), (imgs["edited"].numpy() * 255).astype(np.uint8))
    assert np.array_equal(np.array(Image.open(tmp_path / "exp" / "masks_2" / "mask_0.png")), imgs["mask_scaled"][..., 0].numpy().astype(np.uint8) * 255)
    rd = IO.read_transforms(tmp_path / "exp")
    assert rd["poses"].shape == (2, 4, 4) and np.allclose(rd["poses"][0, :3], c2w.numpy())
    assert rd["fx"] == [10.0, 10.0] and rd["width"] == [10, 10] and rd["reference_indices"] == [0] and rd["num_skipped"] == 0
    assert rd["mask_filenames"][0].name == "mask_0.png" and rd["is_synthetic"] is True and rd["original_scale_factor"] == 0.5
    # DataparserOutputs contract (signerf_dataparser.py:210-228, :301-312): poses as stored, transform / scale from the file
    dp = IO.parse_generated_dataset(tmp_path / "exp", scene_scale=2.0, downscale_factor=2)
    assert len(dp.image_filenames) == 2 and dp.mask_filenames is None            # masks only for merged datasets
    # _get_fname (signerf_dataparser.py:330-357): with downscale_factor 2 the files are the down-sampled copies
    assert [f.relative_to(tmp_path / "exp").as_posix() for f in dp.image_filenames] == ["images_2/image_0.png", "images_2/image_1.png"]
    assert tuple(Image.open(dp.image_filenames[0]).size) == (5, 4)               # == the rescaled camera size below
    assert dp.dataparser_scale == 0.5 and torch.equal(dp.dataparser_transform, torch.eye(4)[:3])
    assert torch.equal(dp.scene_box, torch.tensor([[-2.0, -2.0, -2.0], [2.0, 2.0, 2.0]]))
    assert torch.allclose(dp.cameras.camera_to_worlds[0], c2w) and dp.cameras.fx.tolist() == [5.0, 5.0]
    assert dp.cameras.width.tolist() == [5, 5] and dp.cameras.height.tolist() == [4, 4] and len(dp.cameras) == 2
    assert dp.metadata == {"depth_filenames": None, "depth_unit_scale_factor": 1e-3}
    T["original_indices"] = [1]                                                  # merged dataset: frame 0 gets a white mask
    (tmp_path / "exp" / "transforms.json").write_text(json.dumps(T, indent=4))
    Image.new("L", (10, 8)).save(tmp_path / "exp" / "masks" / "mask_1.png")
    dp = IO.parse_generated_dataset(tmp_path / "exp")
    assert [m.name for m in dp.mask_filenames] == ["white.png", "mask_1.png"]
    assert np.array(Image.open(tmp_path / "exp" / "masks" / "white.png")).min() == 255
    dp2 = IO.parse_generated_dataset(tmp_path / "exp", downscale_factor=2)       # masks follow the images into masks_2/
    assert [m.relative_to(tmp_path / "exp").as_posix() for m in dp2.mask_filenames] == ["masks_2/white.png", "masks_2/mask_1.png"]
    assert (tmp_path / "exp" / "masks_2" / "white.png").exists()
    os.remove(tmp_path / "exp" / "images" / "image_1.png")               # a frame whose image is gone is skipped (:104-107)
    assert IO.read_transforms(tmp_path / "exp" / "transforms.json")["num_skipped"] == 1
    assert len(IO.parse_generated_dataset(tmp_path / "exp").image_filenames) == 1


@pytest.mark.gpu
def test_generate_dataset_end_to_end(tmp_path):
    import signerf_b200.plugin as P
    from oracle import nerfacto_ref as R
    from signerf_b200 import ops
    from tests.helpers import field_from_oracle, ring_cameras

    class Stub(P.Diffuser):
        def diffuse(self, original_image, rendered_image, mask_image=None, condition_image=None):
            return (1.0 - 0.5 * original_image).cpu()

    H, W, ds = 32, 40, 2
    cfg = P.DatasetGeneratorConfig(rows=2, cols=2, width=W, height=H, downscale_factor=ds, mask_dialation=(5, 5), fx=float(W),
                                   fy=float(W), cx=W / 2, cy=H / 2, path=tmp_path, dataset_name="ds")
    gen = cfg.setup(original_transform_matrix=torch.eye(4)[:3], original_scale_factor=1.0,
                    transform_poses_to_original_space=lambda x: x, device="cuda")
    gen.diffuser = Stub(cfg.diffuser, "cuda")
    m = R.make_model(0, dense=True, table_scale=0.5, density_gain=20.0)
    graph = P.FusedNerfactoGraph(field_from_oracle(m), ops.RenderOptions(mode="flat", num_samples=16, mlp_mode=ops.MLP_FP32))
    c2w, _ = ring_cameras(5, W, H)
    gen.generate_dataset(graph, c2w[:3], synthetic_camera_to_worlds=c2w[3:])
    root = tmp_path / "ds"
    meta = json.loads((root / "transforms.json").read_text())
    assert meta["reference_indices"] == [0, 1, 2] and meta["generated_indices"] == [3, 4] and meta["is_synthetic"] is True
    assert len(meta["frames"]) == 5 and meta["frames"][4]["file_path"] == "./images/image_4.png"
    for i in range(5):
        assert Image.open(root / "images" / f"image_{i}.png").size == (W, H)
        assert Image.open(root / f"images_{ds}" / f"image_{i}.png").size == (W // ds, H // ds)
        assert (root / "masks" / f"mask_{i}.png").exists() and (root / "conditions" / f"condition_{i}.png").exists()
    sheet = Image.open(root / "references" / "edited_reference_sheet.png")
    assert sheet.size == (2 * (W // ds), 2 * (H // ds))
    # the rendered tile written to disk is the quantised render of that camera
    cam = P.CameraBatch(c2w[3:4], float(W), float(W), W / 2, H / 2, W, H)
    rgb = gen.render_camera(graph, cam)[0]
    assert np.array_equal(np.array(Image.open(root / "rendered" / "image_3.png")), IO.quantize_image(rgb))
    rd = IO.read_transforms(root)
    assert rd["poses"].shape == (5, 4, 4) and np.allclose(rd["poses"][3, :3], c2w[3].numpy(), atol=1e-6)
    with pytest.raises(ValueError, match="Either original dataset"):
        gen.generate_dataset(graph, c2w[:3])
