"""Blender-synthetic reader (nerfpp_b200/blender.py) against src/load_blender.h:43-217 restated in the test: a scene written to a temp
directory (transforms_*.json + PNG files using all five PNG filter types), then views, intrinsics, near / far and the bounding box."""
import json
import math
import os
import struct
import zlib

import numpy as np
import pytest

from nerfpp_b200 import blender as B


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, img, filters=(0, 1, 2, 3, 4)):
    """8-bit PNG writer that cycles through the given filter types row by row."""
    h, w, c = img.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    rows = img.reshape(h, w * c).astype(np.int32)
    raw = bytearray()
    for y in range(h):
        ft = filters[y % len(filters)]
        cur, prev = rows[y], (rows[y - 1] if y else np.zeros(w * c, dtype=np.int32))
        line = np.zeros(w * c, dtype=np.int32)
        for x in range(w * c):
            a = cur[x - c] if x >= c else 0
            b = prev[x]
            cc = prev[x - c] if x >= c else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, cc))[ft]
            line[x] = (cur[x] - pred) & 255
        raw.append(ft)
        raw.extend(line.astype(np.uint8).tobytes())

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xffffffff)

    comp = zlib.compress(bytes(raw))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + chunk(b"IDAT", comp[:len(comp) // 2]) +
                chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b""))


@pytest.fixture()
def scene(tmp_path):
    rng = np.random.default_rng(0)
    angle = 0.6911112070083618
    counts = {"train": 4, "val": 2, "test": 3}
    imgs = {}
    for split, n in counts.items():
        os.makedirs(tmp_path / split, exist_ok=True)
        frames = []
        for i in range(n):
            pose = B.pose_spherical(-180 + 67.0 * i + (11 if split == "val" else 0), -30.0 + 5 * i, 4.0)
            img = rng.integers(0, 256, size=(12, 16, 4), dtype=np.uint8)
            write_png(tmp_path / split / f"r_{i}.png", img)
            imgs[(split, i)] = img
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.1, "transform_matrix": pose.tolist()})
        with open(tmp_path / f"transforms_{split}.json", "w") as f:
            json.dump({"camera_angle_x": angle, "frames": frames}, f)
    return tmp_path, angle, counts, imgs


def test_png_roundtrip_all_filters_and_channel_counts(tmp_path):
    rng = np.random.default_rng(1)
    for c in (1, 2, 3, 4):
        img = rng.integers(0, 256, size=(11, 7, c), dtype=np.uint8)
        write_png(tmp_path / f"a{c}.png", img)
        assert B.read_png_size(str(tmp_path / f"a{c}.png")) == (11, 7)
        assert np.array_equal(B.read_png(str(tmp_path / f"a{c}.png")), img)
    with open(tmp_path / "bad.png", "wb") as f:
        f.write(b"not a png at all........")
    with pytest.raises(ValueError):
        B.read_png_size(str(tmp_path / "bad.png"))


def test_views_intrinsics_and_splits(scene):
    base, angle, counts, imgs = scene
    d = B.load_blender_data(str(base))                       # testskip: the test split is not read (src/load_blender.h:140-141)
    assert d.SplitsIdx == [4, 2, 0] and len(d.Views) == 6 and [v.ID for v in d.Views] == list(range(6))
    full = B.load_blender_data(str(base), testskip=False)
    assert full.SplitsIdx == [4, 2, 3] and len(full.Views) == 9
    v = d.Views[1]
    assert (v.H, v.W) == (12, 16)
    focal = 0.5 * 16 / math.tan(0.5 * angle)
    assert v.Focal == pytest.approx(focal, rel=1e-6)
    np.testing.assert_allclose(v.K, [[focal, 0, 8], [0, focal, 6], [0, 0, 1]], rtol=1e-6)
    np.testing.assert_array_equal(v.Pose, B.pose_spherical(-180 + 67.0, -25.0, 4.0))
    np.testing.assert_allclose(B.load_image(v), imgs[("train", 1)].astype(np.float32) / 255.0)
    # half_res: H, W and the focal length halve (:162-168); the image is the 2x2 box mean
    half = B.load_blender_data(str(base), half_res=True)
    hv = half.Views[1]
    assert (hv.H, hv.W) == (6, 8) and hv.Focal == pytest.approx(focal / 2, rel=1e-6)
    ref = (imgs[("train", 1)].astype(np.float32) / 255.0).reshape(6, 2, 8, 2, 4).mean(axis=(1, 3))
    np.testing.assert_allclose(B.load_image(hv), ref, rtol=1e-6)


def test_near_far_and_bounding_box(scene):
    base, angle, counts, imgs = scene
    d = B.load_blender_data(str(base))
    origins = np.stack([v.Pose[:3, 3] for v in d.Views[:4]])                 # the training cameras only (:90)
    diag = float(np.linalg.norm(origins.max(0) - origins.min(0)))
    for v in d.Views:
        assert v.Near == pytest.approx(0.15 * diag, rel=1e-6) and v.Far == pytest.approx(0.6 * diag, rel=1e-6)
    # bounding box: near / far points of the four corner rays of every training view (:108-121), rays as GetRays builds them
    pts = []
    for v in d.Views[:4]:
        for px, py in ((0, 0), (15, 0), (0, 11), (15, 11)):
            dirs = np.array([(px - v.K[0, 2]) / v.K[0, 0], -(py - v.K[1, 2]) / v.K[1, 1], -1.0])
            dw = v.Pose[:3, :3].astype(np.float64) @ dirs
            for t in (v.Near, v.Far):
                pts.append(v.Pose[:3, 3] + t * dw)
    pts = np.stack(pts)
    np.testing.assert_allclose(d.BoundingBox, np.concatenate([pts.min(0), pts.max(0)]), rtol=1e-5, atol=1e-5)
    # explicit planes are taken as given (:208-213)
    e = B.load_blender_data(str(base), near=2.0, far=6.0)
    assert all(v.Near == 2.0 and v.Far == 6.0 for v in e.Views)


def test_pose_spherical_and_calibration_helpers():
    p = B.pose_spherical(30.0, -30.0, 4.0)
    assert p.dtype == np.float32 and p.shape == (4, 4)
    np.testing.assert_allclose(np.linalg.norm(p[:3, 3]), 4.0, rtol=1e-6)     # the camera sits on the radius-4 sphere ...
    np.testing.assert_allclose(p[:3, :3] @ p[:3, :3].T, np.eye(3), atol=1e-6)  # ... with an orthonormal frame
    np.testing.assert_allclose(p[:3, 2], p[:3, 3] / 4.0, atol=1e-6)          # looking at the origin (camera looks along -z)
    q = B.pose_spherical(30.0, -30.0, 4.0, 0.5, -0.25, 1.0)
    np.testing.assert_allclose(q[:3, 3] - p[:3, 3], [0.5, -0.25, 1.0], atol=1e-6)
    k = B.get_calibration_matrix(1111.0, 800, 600)
    k2 = B.get_same_fov_calibration_matrix(k, 400, 300)
    np.testing.assert_allclose(k2, [[555.5, 0, 200], [0, 555.5, 150], [0, 0, 1]], rtol=1e-5)
