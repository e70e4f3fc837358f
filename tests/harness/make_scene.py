"""Writes a tiny synthetic Blender-format scene (transforms_{train,val,test}.json + RGBA PNGs) for the harness that
executes the reference's unmodified run_nerf.py.  Two shaded spheres, ray-cast analytically from
cameras on the reference's own spherical path (load/load_blender.py:30-35: radius 4, elevation -30 degrees).
TEST HARNESS ONLY."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "standins"))
import imageio  # noqa: E402  (the harness stand-in: a zlib PNG writer)


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    t = np.eye(4)
    t[2, 3] = radius
    p, th = np.deg2rad(phi_deg), np.deg2rad(theta_deg)
    rp = np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]])
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return flip @ rt @ rp @ t


SPHERES = [(np.array([0.0, 0.0, 0.0]), 0.9, np.array([0.9, 0.25, 0.2])),
           (np.array([0.9, 0.5, 0.4]), 0.45, np.array([0.2, 0.4, 0.9]))]
LIGHT = np.array([0.5, 0.3, 0.8]) / np.linalg.norm([0.5, 0.3, 0.8])


def render(c2w: np.ndarray, H: int, W: int, focal: float) -> np.ndarray:
    i, j = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    dirs = np.stack([(i - 0.5 * W) / focal, -(j - 0.5 * H) / focal, -np.ones_like(i)], -1)
    d = dirs @ c2w[:3, :3].T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = c2w[:3, 3]
    best = np.full((H, W), np.inf)
    rgba = np.zeros((H, W, 4))
    for centre, radius, colour in SPHERES:
        oc = o - centre
        b = (d * oc).sum(-1)
        disc = b * b - (oc @ oc - radius * radius)
        t = -b - np.sqrt(np.maximum(disc, 0))
        hit = (disc > 0) & (t > 0) & (t < best)
        n = (o + t[..., None] * d - centre) / radius
        shade = 0.25 + 0.75 * np.clip((n * LIGHT).sum(-1), 0, 1)
        rgba[hit, :3] = (shade[..., None] * colour)[hit]
        rgba[hit, 3] = 1.0
        best = np.where(hit, t, best)
    return (255 * np.clip(rgba, 0, 1) + 0.5).astype(np.uint8)


def make(root: str, H: int = 64, W: int = 64, n_train: int = 8, n_val: int = 2, n_test: int = 2) -> str:
    angle_x = 0.6911112070083618  # the synthetic Blender scenes' camera_angle_x
    focal = 0.5 * W / np.tan(0.5 * angle_x)
    rs = np.random.RandomState(0)
    for split, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for k in range(n):
            theta = 360.0 * k / n + (0 if split == "train" else 17 + 40 * k) + rs.uniform(-3, 3)
            c2w = pose_spherical(theta, -30.0 + rs.uniform(-8, 8), 4.0)
            imageio.imwrite(os.path.join(root, split, f"r_{k}.png"), render(c2w, H, W, focal))
            frames.append({"file_path": f"./{split}/r_{k}", "transform_matrix": c2w.tolist()})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as fh:
            json.dump({"camera_angle_x": angle_x, "frames": frames}, fh)
    return root


if __name__ == "__main__":
    print(make(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_scene")))
