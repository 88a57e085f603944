"""Rendering front-end of frame synthesis: batched stickman rasterisation on the GPU.

The reference's video-rendering loop (``data/data_conversions_3d.py:1130-1173``) rasterises ONE stickman per frame on
the host with ``make_joint_img`` (``lib/utils.py:325-513``: ``cv2.fillPoly`` + ``cv2.line`` into three uint8 planes),
applies ``ToTensor()`` + ``x * 2 - 1``, copies it to the GPU and calls ``transfer`` with batch 1.
:class:`StickmanRasterizer` draws a whole batch of poses in one launch (``vub_stickman_render``), bit-exactly, as the
uint8 HWC images that the stem kernel ingests directly (``vub_stem_fwd_u8``) — so a sequence goes from projected 2-D
joints to video frames without touching the host: ``PoseRenderer(vunet, joint_model)(appearance_u8, joints)``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _native as N
from . import ops


def get_line_colors(n_lines_per_channel: Sequence[int]):
    """lib/utils.py:51-66 (the ``diff_line_colors`` option of the datasets)."""
    line_colors = []
    for channel, nr_lines in enumerate(n_lines_per_channel):
        interval = int(255 // (nr_lines + 1))
        cols = []
        for i in range(nr_lines):
            col = [0, 0, 0]
            col[channel] = (i + 1) * interval
            cols.append(col)
        line_colors.append(cols)
    return line_colors


def project_poses(poses3d: torch.Tensor, extrinsics: torch.Tensor, intrinsics, image_size, target_size: int) -> torch.Tensor:
    """3-D joints ``[B, J, 3]`` -> 2-D joints ``[B, J, 2]`` in the pixel coordinates of a ``target_size`` stickman: the
    per-frame prologue of the reference's rendering loop (``apply_affine_transform`` with the ``[R|t]`` extrinsics,
    ``camera_projection`` with ``(f_x, x_0, f_y, y_0)``, rescaling by ``target_size / image_size``;
    ``data/data_conversions_3d.py:588-605, 892-912, 1131-1132, 1150-1152``) for a whole batch, in float64 like the
    numpy original.  Plain tensor algebra on a few kilobytes (plumbing, not a kernel of the path)."""
    x = poses3d.to(torch.float64)
    M = torch.as_tensor(extrinsics, dtype=torch.float64, device=x.device)
    fx, x0, fy, y0 = (float(v) for v in intrinsics)
    cam = torch.tensor([[fx, 0.0, x0], [0.0, fy, y0], [0.0, 0.0, 1.0]], dtype=torch.float64, device=x.device)
    x_hom = torch.cat([x, torch.ones_like(x[..., :1])], dim=-1)
    pose_c = x_hom @ M.T
    pose_i = ((pose_c / pose_c[..., -1:]) @ cam.T)[..., :-1]
    size = torch.as_tensor(image_size, dtype=torch.float64, device=x.device).reshape(-1, 1, 2)
    return pose_i * (float(target_size) / size)


class StickmanRasterizer:
    """``imgs = StickmanRasterizer(joint_model, size)(joints)`` == ``np.stack([make_joint_img([size, size, 3], j,
    joint_model, line_colors) for j in joints])`` of the reference, as a uint8 ``[B, size, size, 3]`` CUDA tensor.

    ``joint_model``: any object with the reference ``JointModel``'s fields ``body, right_lines, left_lines, head_lines,
    face, rshoulder, lshoulder, headup`` (``lib/utils.py``); ``line_colors``: ``None`` or the nested list
    ``get_line_colors`` returns; ``joints``: ``[B, J, 2]`` image coordinates (x, y), float64 like the reference's numpy
    arrays (the integer pixel is the C truncation of the double), negative = missing.  Line thickness is 1, which is
    what every shipped config produces (``stickman_scale`` unset -> ``scale_factor=None``)."""

    def __init__(self, joint_model, img_size: int, line_colors=None, scale_factor=None):
        if scale_factor is not None and int(img_size // scale_factor) != 1:
            raise NotImplementedError("StickmanRasterizer draws lines of thickness 1 (no shipped config sets stickman_scale)")
        self.size = int(img_size)
        if (self.size * self.size * 3) % 16:
            raise ValueError("img_size**2 * 3 must be a multiple of 16")
        m = N.StickmanModel()

        def put(n_field, arr_field, lines, cap):
            lines = list(lines)
            if len(lines) > cap:
                raise NotImplementedError(f"joint model has {len(lines)} {arr_field} lines (max {cap})")
            setattr(m, n_field, len(lines))
            arr = getattr(m, arr_field)
            for i, (a, b) in enumerate(lines):
                arr[i][0], arr[i][1] = int(a), int(b)

        body = list(joint_model.body)
        if len(body) > 8:
            raise NotImplementedError("joint model body polygon has more than 8 joints")
        m.n_body = len(body)
        for i, j in enumerate(body):
            m.body[i] = int(j)
        put("n_right", "right", joint_model.right_lines, 16)
        put("n_left", "left", joint_model.left_lines, 16)
        put("n_head", "head", joint_model.head_lines, 8)
        put("n_face", "face", joint_model.face, 16)
        m.rshoulder, m.lshoulder, m.headup = int(joint_model.rshoulder), int(joint_model.lshoulder), int(joint_model.headup)
        m.colored = 0 if line_colors is None else 1
        if line_colors is not None:
            def colors(chan_field, col_field, table, n):
                for i in range(n):
                    nz = [k for k, v in enumerate(table[i]) if v != 0]
                    if len(nz) != 1:
                        raise ValueError("line_colors entries must have exactly one non-zero channel (get_line_colors)")
                    getattr(m, chan_field)[i] = nz[0]
                    getattr(m, col_field)[i] = int(table[i][nz[0]])
            colors("right_chan", "right_col", line_colors[0], m.n_right)
            colors("left_chan", "left_col", line_colors[1], m.n_left)
            colors("head_chan", "head_col", line_colors[2], max(m.n_head, m.n_face))
        self._model = m
        self.n_joints_min = 1 + max([m.rshoulder, m.lshoulder, m.headup] + body
                                    + [j for l in (joint_model.right_lines, joint_model.left_lines, joint_model.head_lines,
                                                   joint_model.face) for ab in l for j in ab])

    def __call__(self, joints: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not joints.is_cuda:
            raise RuntimeError("vunet_b200 has no CPU path: joints must live on a CUDA device")
        if joints.dim() != 3 or joints.shape[2] != 2 or joints.shape[1] < self.n_joints_min:
            raise ValueError(f"joints must be [B, J >= {self.n_joints_min}, 2], got {tuple(joints.shape)}")
        joints = joints.to(torch.float64).contiguous()
        B = joints.shape[0]
        s = self.size
        if out is None:
            out = torch.empty((B, s, s, 3), dtype=torch.uint8, device=joints.device)
        elif out.dtype != torch.uint8 or tuple(out.shape) != (B, s, s, 3) or not out.is_contiguous():
            raise ValueError("out must be a contiguous uint8 [B, size, size, 3] tensor")
        N.check(ops._launch("stickman", N.lib().vub_stickman_render, joints.data_ptr(), B, joints.shape[1],
                            C.byref(self._model), s, s, out.data_ptr(), ops._stream(),
                            bytes_=float(out.numel() + joints.numel() * 8)), "vub_stickman_render")
        return out


class PoseRenderer:
    """Joints in, video frames out: ``frames = PoseRenderer(vunet, joint_model)(appearance, joints)`` — stickman
    rasterisation, ``ToTensor()`` + ``x * 2 - 1``, ``transfer`` and the uint8 frame conversion of the reference's
    rendering loop (``data/data_conversions_3d.py:1154-1173``) with nothing but kernels of this package in between.

    ``appearance``: one image per frame, uint8 HWC ``[B, Sx, Sx, Cx]`` or fp32 NCHW in [-1, 1]; ``joints``: ``[B, J, 2]``
    already scaled to the model's ``spatial_size`` (the loop's ``pose_i_rescaled``).  Returns uint8 ``[B, S, S, 3]``."""

    def __init__(self, model, joint_model, line_colors=None):
        self.model = model.eval()
        self.raster = StickmanRasterizer(joint_model, model.spatial_size, line_colors)

    @torch.no_grad()
    def __call__(self, appearance: torch.Tensor, joints: torch.Tensor, uint8_frames: bool = True) -> torch.Tensor:
        stick = self.raster(joints)
        return self.model.transfer(appearance, stick, uint8_frames=uint8_frames)


class BehaviorVideoRenderer:
    """BASELINE.json configs[4] as one object: noise -> behavior-net pose sequences (``PoseSequenceSampler``: latent-flow
    reverse pass + residual LSTM decoder, one CUDA graph) -> camera projection -> stickman rasterisation -> VUnet
    synthesis with the appearance encoded once per sequence (``SequenceSynthesizer``, one CUDA graph) -> uint8 frames.
    The reference does this per frame on the host (lib/logging.py:280-300 for the poses,
    data/data_conversions_3d.py:1130-1173 for the rendering); here nothing but the frames has to leave the GPU.

    ``n_sampled`` sequences are drawn per sampler replay (256 is the efficient batch); every sequence is rendered by one
    synthesizer replay.  ``joint_model`` indexes the ``n_kps / 3`` generated joints; ``extrinsics`` (3x4), ``intrinsics``
    (f_x, x_0, f_y, y_0) and ``image_size`` as for :func:`project_poses`.  Sequences are independent: shard them over
    GPUs with ``synthesis.shard_range`` (no collective)."""

    def __init__(self, vunet, flow, net, joint_model, extrinsics, intrinsics, image_size, frames: int, n_sampled: int = 256,
                 line_colors=None, use_graph: bool = True):
        from .behavior import PoseSequenceSampler
        from .synthesis import SequenceSynthesizer
        if net.n_kps % 3:
            raise ValueError("BehaviorVideoRenderer: the behavior net must generate 3-D joints (n_kps a multiple of 3)")
        self.frames, self.n_sampled = frames, n_sampled
        self.sampler = PoseSequenceSampler(flow, net, n_sampled, frames, use_graph=use_graph)
        self.raster = StickmanRasterizer(joint_model, vunet.spatial_size, line_colors)
        self.synth = SequenceSynthesizer(vunet, 1, frames, use_graph=use_graph, uint8_frames=True, uint8_inputs=True)
        self.camera = (extrinsics, intrinsics, image_size)
        self.size = vunet.spatial_size
        self.n_joints = net.n_kps // 3

    def sample_poses(self, z: torch.Tensor, x_start: torch.Tensor) -> torch.Tensor:
        """[n_sampled, frames, J, 2] joints in stickman pixel coordinates for noise ``z`` and start poses ``x_start``."""
        poses = self.sampler(z, x_start)
        j3 = poses.reshape(self.n_sampled * self.frames, self.n_joints, 3)
        j2 = project_poses(j3, self.camera[0], self.camera[1], self.camera[2], self.size)
        return j2.reshape(self.n_sampled, self.frames, self.n_joints, 2)

    def render_sequence(self, appearance_u8: torch.Tensor, joints: torch.Tensor, out: Optional[torch.Tensor] = None):
        """One sequence: ``appearance_u8`` [1, Sx, Sx, Cx] uint8, ``joints`` [frames, J, 2] -> uint8 frames
        [frames, S, S, 3] (the synthesizer's static output, or copied into ``out``, e.g. pinned host memory)."""
        self.raster(joints, out=self.synth.c)        # stickmen straight into the graph's static input
        self.synth.x.copy_(appearance_u8, non_blocking=True)
        frames = self.synth.run()
        if out is not None:
            out.copy_(frames, non_blocking=True)
            return out
        return frames
