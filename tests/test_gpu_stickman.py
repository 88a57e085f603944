"""GPU (-m gpu): the batched stickman rasteriser (vub_stickman_render) — bit-exact with the outputs of the unmodified
reference function (tests/golden/stickman.npz) and with the CPU oracle on random poses incl. missing joints and limbs
and bodies leaving the image on every side; joints -> frames end to end."""
import contextlib
import io

import numpy as np
import pytest
import torch

import stickman_oracle as S
from test_stickman import MODELS, golden_cases

pytestmark = pytest.mark.gpu


def test_rasterizer_matches_reference_golden_bit_for_bit():
    import bdvs_b200 as B
    n = 0
    for key, jm, lc, size, joints, imgs in golden_cases():
        r = B.StickmanRasterizer(jm, size, lc)
        got = r(torch.from_numpy(joints).cuda()).cpu().numpy()
        assert got.dtype == np.uint8 and got.shape == imgs.shape
        assert np.array_equal(got, imgs), (key, int((got != imgs).sum()))
        n += len(joints)
    assert n == 144


def test_rasterizer_matches_oracle_on_random_poses():
    import bdvs_b200 as B
    rng = np.random.default_rng(5)
    for name, jm in MODELS.items():
        for size in (64, 256):
            lc = S.get_line_colors((5, 5, len(jm.head_lines)))
            joints = rng.normal(size / 2, size * 0.4, size=(48, 32, 2))     # many limbs leave the image
            joints[rng.random((48, 32)) < 0.1] = -3.0                        # missing joints
            joints[0, :, :] = -1.0                                            # a frame with nothing to draw
            r = B.StickmanRasterizer(jm, size, lc)
            got = r(torch.from_numpy(joints).cuda()).cpu().numpy()
            for b in range(joints.shape[0]):
                want = S.make_joint_img([size, size, 3], joints[b], jm, line_colors=lc)
                assert np.array_equal(got[b], want), (name, size, b)
            assert got[0].max() == 0


def test_pose_renderer_equals_reference_pipeline():
    """joints -> stickman -> ToTensor / x*2-1 -> transfer -> uint8 frame, against the same steps done the reference's
    way around the module API (oracle stickman on the host, fp32 conversion, transfer, scale_img * 255)."""
    import bdvs_b200 as B
    import vunet_oracle as O
    cfg = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0, nf_start=32, nf_max=128,
               spatial_size=128, bottleneck_factor=2, box_factor=2, n_channels_x=3)
    with contextlib.redirect_stdout(io.StringIO()):
        m = B.VunetAlter(**cfg)
    m.load_state_dict(O.det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
    m = m.cuda().eval()
    jm = MODELS["h36m_full"]
    rng = np.random.default_rng(11)
    joints = rng.normal(64, 22, size=(3, 32, 2))
    app = torch.randint(0, 256, (3, 128, 128, 3), dtype=torch.uint8)
    torch.manual_seed(3)
    frames = B.PoseRenderer(m, jm)(app.cuda(), torch.from_numpy(joints).cuda())
    stick = np.stack([S.make_joint_img([128, 128, 3], j, jm) for j in joints])
    to_f = lambda t: ((t.float().div(255) * 2.0) - 1.0).permute(0, 3, 1, 2).contiguous()  # noqa: E731
    torch.manual_seed(3)
    with torch.no_grad():
        img = m.transfer(to_f(app).cuda(), to_f(torch.from_numpy(stick)).cuda())
    want = (torch.clamp((img + 1.0) / 2.0, 0.0, 1.0) * 255.0).permute(0, 2, 3, 1).cpu().numpy().astype("uint8")
    assert frames.dtype == torch.uint8 and np.array_equal(frames.cpu().numpy(), want)
