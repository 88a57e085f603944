"""Pins oracle/stickman_oracle.py to the UNMODIFIED reference rasteriser and writes tests/golden/stickman.npz.

Run in the build container (needs /root/reference and cv2):  python oracle/gen_stickman_golden.py
The reference module ``lib/utils.py`` cannot be imported (kornia, matplotlib, ... are absent), so the SOURCE TEXT of its
``make_joint_img`` (lines 325-513) is extracted with ``ast`` and executed unchanged against the real ``cv2`` / ``numpy``.
Two shims for this image's numpy 2.x, neither touching the algorithm: ``np.float`` (removed alias of ``float``) and
``np.int_`` wrapped to return int32 (numpy's default integer is 64-bit here, which this cv2 build refuses as point
type; the pinned environment's values are identical, the coordinates are < 2^31).
"""
import ast
import os
import sys
from collections import namedtuple

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import stickman_oracle as S  # noqa: E402

REF = "/root/reference/lib/utils.py"


def load_reference():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "make_joint_img")
    text = "\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])

    class _NP:  # numpy with the two shims
        def __getattr__(self, k):
            return getattr(np, k)

        float = float

        @staticmethod
        def int_(a):
            return np.asarray(a).astype(np.int32)

    JointModel = namedtuple("JointModel", "body right_lines left_lines head_lines face rshoulder lshoulder headup")
    ns = {"np": _NP(), "cv2": cv2, "JointModel": JointModel}
    exec(compile(text, REF, "exec"), ns)
    return ns["make_joint_img"], (fn.lineno, fn.end_lineno)


# the two Human3.6m joint models of data/human36m.py:77-86 and 136-152 (32-joint skeleton)
MODELS = {
    "h36m_small": S.JointModel(body=[25, 17, 6, 1], right_lines=[(3, 2), (2, 1), (1, 25), (25, 26), (26, 30)],
                               left_lines=[(8, 7), (7, 6), (6, 17), (17, 18), (18, 22)], head_lines=[], face=[],
                               rshoulder=25, lshoulder=17, headup=15),
    "h36m_full": S.JointModel(body=[1, 25, 13, 17, 6], right_lines=[(3, 2), (2, 1), (1, 25), (25, 26), (26, 27)],
                              left_lines=[(8, 7), (7, 6), (6, 17), (17, 18), (18, 19)], head_lines=[(13, 14), (14, 15)],
                              face=[], rshoulder=25, lshoulder=17, headup=15),
}


def random_pose(rng, size, spread):
    """32 joints around the image centre; a few joints missing (negative) or outside the image below / right."""
    j = rng.normal(loc=size / 2, scale=size * spread, size=(32, 2))
    miss = rng.random(32) < 0.08
    j[miss] = -1.0
    return j.astype(np.float64)


def main():
    ref_fn, lines = load_reference()
    rng = np.random.default_rng(20261017)
    out = {}
    n_checked = 0
    for name, jm in MODELS.items():
        for colored in (False, True):
            lc = S.get_line_colors((len(jm.left_lines), len(jm.right_lines), len(jm.head_lines) + len(jm.face))) \
                if colored else None
            for size, n, spread in ((64, 24, 0.22), (256, 6, 0.2), (128, 6, 0.35)):
                joints, imgs = [], []
                for _ in range(n):
                    j = random_pose(rng, size, spread)
                    want = ref_fn([size, size, 3], j, jm, line_colors=lc, scale_factor=None)
                    got = S.make_joint_img([size, size, 3], j, jm, line_colors=lc)
                    assert want.dtype == np.uint8 and want.shape == (size, size, 3)
                    assert np.array_equal(want, got), (name, colored, size, int((want != got).sum()))
                    n_checked += 1
                    joints.append(j)
                    imgs.append(want)
                key = f"{name}_{'colored' if colored else 'plain'}_{size}"
                out[key + "_joints"] = np.stack(joints)
                out[key + "_imgs"] = np.stack(imgs)
    path = os.path.join(ROOT, "tests", "golden", "stickman.npz")
    np.savez_compressed(path, **out)
    print(f"reference make_joint_img (lib/utils.py:{lines[0]}-{lines[1]}) == oracle on {n_checked} poses; wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KB), cv2 {cv2.__version__}, numpy {np.__version__}")


if __name__ == "__main__":
    main()
