"""CPU: the stickman oracle (oracle/stickman_oracle.py) against the committed outputs of the UNMODIFIED reference
rasteriser (tests/golden/stickman.npz, written by oracle/gen_stickman_golden.py from lib/utils.py:325-513 + cv2), and
its two OpenCV primitives against cv2 itself where cv2 is importable."""
import os
import random

import numpy as np
import pytest

import stickman_oracle as S
from conftest import GOLDEN_DIR

MODELS = {
    "h36m_small": S.JointModel(body=[25, 17, 6, 1], right_lines=[(3, 2), (2, 1), (1, 25), (25, 26), (26, 30)],
                               left_lines=[(8, 7), (7, 6), (6, 17), (17, 18), (18, 22)], head_lines=[], face=[],
                               rshoulder=25, lshoulder=17, headup=15),
    "h36m_full": S.JointModel(body=[1, 25, 13, 17, 6], right_lines=[(3, 2), (2, 1), (1, 25), (25, 26), (26, 27)],
                              left_lines=[(8, 7), (7, 6), (6, 17), (17, 18), (18, 19)], head_lines=[(13, 14), (14, 15)],
                              face=[], rshoulder=25, lshoulder=17, headup=15),
}


def golden_cases():
    fx = np.load(os.path.join(GOLDEN_DIR, "stickman.npz"))
    for key in sorted(k[:-7] for k in fx.files if k.endswith("_joints")):
        name, kind, size = key.rsplit("_", 2)
        jm = MODELS[name]
        lc = S.get_line_colors((len(jm.left_lines), len(jm.right_lines), len(jm.head_lines) + len(jm.face))) \
            if kind == "colored" else None
        yield key, jm, lc, int(size), fx[key + "_joints"], fx[key + "_imgs"]


def test_oracle_reproduces_reference_make_joint_img():
    n = 0
    for key, jm, lc, size, joints, imgs in golden_cases():
        for j, want in zip(joints, imgs):
            got = S.make_joint_img([size, size, 3], j, jm, line_colors=lc)
            assert np.array_equal(got, want), key
            n += 1
    assert n == 144
    # the fixtures are not trivial: bodies, limbs and throats are present
    _, _, _, _, _, imgs = next(golden_cases())
    assert (imgs[..., 2] == 255).any() and (imgs[..., 1] == 127).any() and (imgs[..., 0] > 0).any()


def test_line_and_polygon_primitives_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rnd = random.Random(7)
    for _ in range(4000):  # lines, end points up to 30 pixels outside the image (clipLine + Bresenham)
        a = (rnd.randint(-30, 93), rnd.randint(-30, 93))
        b = (rnd.randint(-30, 93), rnd.randint(-30, 93))
        want = np.zeros((64, 64), np.uint8)
        cv2.line(want, a, b, color=200, thickness=1)
        got = np.zeros((64, 64), np.uint8)
        S.draw_line(got, a, b, 200)
        assert np.array_equal(got, want), (a, b)
    for _ in range(6000):  # polygons with 3-6 vertices anywhere up to 40 pixels outside the image
        pts = [(rnd.randint(-40, 110), rnd.randint(-40, 110)) for _ in range(rnd.choice([3, 4, 5, 6]))]
        want = np.zeros((64, 64), np.uint8)
        cv2.fillPoly(want, np.array([pts], dtype=np.int32), 127)
        got = np.zeros((64, 64), np.uint8)
        S.fill_poly(got, pts, 127)
        assert np.array_equal(got, want), pts


def test_rasterizer_host_side_contract():
    import bdvs_b200 as B
    import torch
    from bdvs_b200 import _native
    import ctypes
    jm = MODELS["h36m_full"]
    r = B.StickmanRasterizer(jm, 256, B.get_line_colors((5, 5, 2)))
    assert r.n_joints_min == 28 and r._model.n_head == 2 and list(r._model.head_col[:2]) == [85, 170]
    assert B.get_line_colors((5, 5, 2)) == S.get_line_colors((5, 5, 2))
    assert ctypes.sizeof(_native.StickmanModel) == 4 * (1 + 8 + 1 + 32 + 1 + 32 + 1 + 16 + 1 + 32 + 3 + 1 + 6 * 16)
    with pytest.raises(NotImplementedError):
        B.StickmanRasterizer(jm, 256, scale_factor=64)  # thickness 4
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(torch.zeros(1, 32, 2, dtype=torch.float64))


def test_project_poses_matches_reference_formulas():
    """data/data_conversions_3d.py:588-605 (apply_affine_transform), 892-912 (camera_projection), 1150-1152 (rescaling)
    restated with numpy exactly as the reference writes them."""
    import bdvs_b200 as B
    import torch
    rng = np.random.default_rng(3)
    poses = rng.normal(size=(5, 17, 3)) + np.array([0.0, 0.0, 5.0])
    M = np.concatenate([np.linalg.qr(rng.normal(size=(3, 3)))[0], rng.normal(size=(3, 1))], axis=1)
    M[2, 3] += 6.0
    intr = (1145.0, 512.5, 1143.8, 515.4)
    sizes = np.array([[1000.0, 1002.0]] * 5)
    want = []
    for p, s in zip(poses, sizes):
        x_hom = np.concatenate([p, np.ones((p.shape[0], 1), dtype=p.dtype)], axis=-1)
        pose_c = x_hom @ M.T
        cam = np.asarray([[intr[0], 0.0, intr[1]], [0.0, intr[2], intr[3]], [0.0, 0.0, 1.0]])
        pose_i = ((pose_c / np.expand_dims(pose_c[..., -1], axis=-1)) @ cam.T)[..., :-1]
        want.append(pose_i * (np.full((1, 2), 256, dtype=float) / np.expand_dims(s, axis=0)))
    got = B.project_poses(torch.from_numpy(poses), torch.from_numpy(M), intr, torch.from_numpy(sizes), 256).numpy()
    assert got.dtype == np.float64 and np.allclose(got, np.stack(want), rtol=1e-13, atol=1e-10)
