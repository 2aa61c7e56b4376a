"""Blender-synthetic dataset reader: the host-side mirror of src/load_blender.h, so that a scene the reference trains on can be
fed to this engine with the same views, intrinsics, near / far planes and bounding box.

  load_blender_data(basedir, near, far, half_res, testskip)   src/load_blender.h:127-217   -> DatasetParams (Views, Splits, SplitsIdx, BoundingBox)
  get_bounds_for_obj / get_bbox3d_for_obj                       :84-125
  pose_spherical, get_calibration_matrix, get_same_fov_calibration_matrix   :43-82

The reference reads the images with OpenCV (cv::imread IMREAD_UNCHANGED, :155) and nlohmann::json; neither is in this image, so the JSON side
uses the standard library and the PNG side is a small decoder of its own (read_png: 8-bit gray / RGB / RGBA, non-interlaced — what Blender
writes).  Only the image SIZE is needed to build a view (read_png_size reads the IHDR chunk); pixels are decoded on demand (load_image).
Host code only: nothing here touches the GPU; HashNeRF.set_camera takes the decoded image and (K, Pose) of a view."""
from __future__ import annotations

import json
import math
import os
import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

SPLITS = ("train", "val", "test")          # NeRFDatasetParams::Splits (src/NeRFDataset.h)


def _rot(rows):
    return np.array(rows, dtype=np.float32)


def pose_spherical(theta: float, phi: float, radius: float, x: float = 0.0, y: float = 0.0, z: float = 0.0) -> np.ndarray:
    """src/load_blender.h:43-57: camera-to-world of a camera on a sphere, angles in degrees, fp32 like the reference."""
    ph, th = np.float32(phi / 180.0 * math.pi), np.float32(theta / 180.0 * math.pi)
    c2w = _rot([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]])
    c2w = _rot([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]]) @ c2w
    c2w = _rot([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]]) @ c2w
    c2w = _rot([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    c2w[0, 3] += x
    c2w[1, 3] += y
    c2w[2, 3] += z
    return c2w


def get_calibration_matrix(focal: float, w: float, h: float) -> np.ndarray:
    """src/load_blender.h:60-66."""
    return np.array([[focal, 0, 0.5 * w], [0, focal, 0.5 * h], [0, 0, 1]], dtype=np.float32)


def get_same_fov_calibration_matrix(k: np.ndarray, new_w: float, new_h: float) -> np.ndarray:
    """src/load_blender.h:69-81: new image size, same field of view along the longer side."""
    focal, w, h = float(k[0, 0]), float(k[0, 2]) * 2, float(k[1, 2]) * 2
    angle = 2.0 * math.atan(max(w, h) / 2 / focal)
    return get_calibration_matrix(0.5 * max(new_w, new_h) / math.tan(0.5 * angle), new_w, new_h)


@dataclass
class View:
    """One frame of a transforms_*.json (View, src/NeRFDataset.h)."""
    ID: int
    ImagePath: str
    H: int
    W: int
    Focal: float
    K: np.ndarray          # [3,3] fp32
    Pose: np.ndarray       # [4,4] fp32 camera-to-world
    Near: float = 0.0
    Far: float = 0.0
    HalfRes: bool = False


@dataclass
class DatasetParams:
    Views: list = field(default_factory=list)
    Splits: tuple = SPLITS
    SplitsIdx: list = field(default_factory=lambda: [0, 0, 0])     # number of views per split, in Splits order (train views come first)
    BoundingBox: np.ndarray | None = None                            # [6] = (min xyz, max xyz)


# ------------------------------------------------------------------------------------------------------------ PNG
_PNG_MAGIC = b"\x89PNG\r\n\x1a\n"
_CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}      # colour type -> samples per pixel (3 = palette: not written by Blender, not read here)


def read_png_size(path: str) -> tuple[int, int]:
    """(height, width) from the IHDR chunk."""
    with open(path, "rb") as f:
        head = f.read(24)
    if head[:8] != _PNG_MAGIC or head[12:16] != b"IHDR":
        raise ValueError(f"{path}: not a PNG file")
    w, h = struct.unpack(">II", head[16:24])
    return h, w


def read_png(path: str) -> np.ndarray:
    """uint8 [H, W, C] with all channels kept (cv::IMREAD_UNCHANGED, src/load_blender.h:155; channel order R,G,B,A — OpenCV's is B,G,R,A).
    8-bit gray / gray+alpha / RGB / RGBA, non-interlaced; anything else raises."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != _PNG_MAGIC:
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, ihdr = 8, [], None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    if ihdr is None:
        raise ValueError(f"{path}: no IHDR chunk")
    w, h, depth, ctype, _, _, interlace = ihdr
    if depth != 8 or ctype not in _CHANNELS or interlace != 0:
        raise ValueError(f"{path}: only 8-bit non-interlaced gray / RGB / RGBA PNGs are read (bit depth {depth}, colour type {ctype}, interlace {interlace})")
    c = _CHANNELS[ctype]
    stride = w * c
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8).reshape(h, stride + 1)
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:                                        # Up
            cur = (line + prev) & 255
        elif ft == 1:                                        # Sub: a running sum per channel
            cur = np.cumsum(line.reshape(w, c), axis=0).reshape(-1) & 255
        else:                                                # Average / Paeth depend on the reconstructed left neighbour: per pixel
            cur = np.zeros(stride, dtype=np.int32)
            for x in range(stride):
                a = cur[x - c] if x >= c else 0
                b = prev[x]
                if ft == 3:
                    pred = (a + b) >> 1
                elif ft == 4:
                    cc = prev[x - c] if x >= c else 0
                    p = a + b - cc
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - cc)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
                else:
                    raise ValueError(f"{path}: bad filter type {ft}")
                cur[x] = (line[x] + pred) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, c)


def load_image(view: View) -> np.ndarray:
    """float32 [H, W, C] in [0, 1] at the view's resolution (half_res: cv::resize INTER_LINEAR by exactly 1/2 = the 2x2 box mean)."""
    img = read_png(view.ImagePath).astype(np.float32) / 255.0
    if view.HalfRes:
        h2, w2 = img.shape[0] // 2, img.shape[1] // 2
        img = img[:h2 * 2, :w2 * 2].reshape(h2, 2, w2, 2, -1).mean(axis=(1, 3))
    return img


# ------------------------------------------------------------------------------------------------------------ bounds
def _corner_ray(view: View, px: int, py: int):
    """GetRays (src/RayUtils.h:23-46) for one pixel: same expression as csrc/rays.cu pixel_ray."""
    k, c2w = view.K, view.Pose
    d_cam = np.array([(px - k[0, 2]) / k[0, 0], -(py - k[1, 2]) / k[1, 1], -1.0], dtype=np.float32)
    return c2w[:3, 3].astype(np.float32), (c2w[:3, :3] @ d_cam).astype(np.float32)


def get_bounds_for_obj(data: DatasetParams) -> tuple[float, float]:
    """src/load_blender.h:84-97: near / far = 0.15 / 0.6 of the diagonal of the training cameras' bounding box."""
    origins = np.stack([v.Pose[:3, 3] for v in data.Views[:data.SplitsIdx[0]]]).astype(np.float32)
    d = float(np.linalg.norm(np.maximum(origins.max(0), -1e8) - np.minimum(origins.min(0), 1e8)))
    return 0.15 * d, 0.6 * d


def get_bbox3d_for_obj(data: DatasetParams) -> np.ndarray:
    """src/load_blender.h:100-125: box of the near and far points of the four corner rays of every training view."""
    lo, hi = np.full(3, 1e8, dtype=np.float32), np.full(3, -1e8, dtype=np.float32)
    for v in data.Views[:data.SplitsIdx[0]]:
        for px, py in ((0, 0), (v.W - 1, 0), (0, v.H - 1), (v.W - 1, v.H - 1)):
            o, d = _corner_ray(v, px, py)
            for t in (v.Near, v.Far):
                p = o + np.float32(t) * d
                lo, hi = np.minimum(lo, p), np.maximum(hi, p)
    return np.concatenate([lo, hi])


def load_blender_data(basedir: str, near: float = 0.0, far: float = 0.0, half_res: bool = False, testskip: bool = True) -> DatasetParams:
    """src/load_blender.h:127-217.  testskip drops the whole test split (:140-141), as the reference does."""
    result = DatasetParams()
    for i_split, split in enumerate(result.Splits):
        if testskip and split == "test":
            continue
        with open(os.path.join(basedir, f"transforms_{split}.json")) as f:
            meta = json.load(f)
        angle = float(meta["camera_angle_x"])
        for frame in meta["frames"]:
            path = os.path.join(basedir, frame["file_path"] + ".png")
            h, w = read_png_size(path)
            focal = 0.5 * w / math.tan(0.5 * angle)
            if half_res:
                h, w, focal = h // 2, w // 2, focal / 2
            pose = np.zeros((4, 4), dtype=np.float32)
            rows = frame["transform_matrix"]
            for r, row in enumerate(rows):
                pose[r, :len(row)] = row
            result.SplitsIdx[i_split] += 1
            result.Views.append(View(ID=len(result.Views), ImagePath=path, H=h, W=w, Focal=focal, K=get_calibration_matrix(focal, w, h), Pose=pose,
                                     HalfRes=half_res))
    bounds = get_bounds_for_obj(result) if (near == 0.0 or far == 0.0) else (0.0, 0.0)
    for v in result.Views:
        v.Near = bounds[0] if near == 0 else near
        v.Far = bounds[1] if far == 0 else far
    result.BoundingBox = get_bbox3d_for_obj(result)
    return result
