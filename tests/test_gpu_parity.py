"""GPU (-m gpu): the CUDA path, called through the module API / torch ops / C ABI, against the CPU
oracle on the same hashed inputs, against the committed golden vectors generated from the reference,
and through size-independent properties at the BASELINE.json sizes.

Tolerance (BASELINE.json north_star): relative L2 <= 1e-2 for bf16 images / latents; latent noise
bit-exact with torch's generator.
"""
import contextlib
import io

import pytest
import torch

import vunet_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2

ARCH = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0, nf_start=32, nf_max=128)
H256 = dict(ARCH, spatial_size=256, bottleneck_factor=2, box_factor=2, n_channels_x=3)
D256 = dict(ARCH, spatial_size=256, bottleneck_factor=2, box_factor=2, n_channels_x=30)
M128 = dict(ARCH, spatial_size=128, bottleneck_factor=1, box_factor=1, n_channels_x=30)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(cfg, sd=None):
    import bdvs_b200 as B
    cfg = {k: v for k, v in cfg.items() if k != "batch"}
    with contextlib.redirect_stdout(io.StringIO()):
        m = B.VunetAlter(**cfg)
    if sd is None:
        sd = O.det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def cuda(ts):
    return [t.cuda() for t in ts]


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny3", "tiny30", "market128"])
def test_model_matches_reference_golden(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    x, c, eps, prior = O.make_inputs(cfg, cfg["batch"])
    with torch.no_grad():
        imgs, means, logstds, acts = m(x.cuda(), c.cuda(), eps=cuda(eps))
        timgs = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps))
    assert imgs.shape == fx["imgs"].shape and imgs.dtype == torch.float32
    assert rel(imgs, fx["imgs"]) < BF16_TOL
    assert rel(timgs, fx["transfer"]) < BF16_TOL
    for i in range(2):
        assert rel(means[i], fx["means"][i]) < BF16_TOL
        assert rel(logstds[i], fx["logstds"][i]) < BF16_TOL
    # return structure of the reference: activations = (hs left after ed's pops, means, logstds)
    assert len(acts[0]) == fx["n_hs_left"]
    assert [tuple(a.shape) for a in acts[0]] == fx["hs_left_shapes"]
    for a, ref_mean_abs in zip(acts[0], fx["hs_left_mean_abs"]):
        assert abs(float(a.float().abs().mean()) - ref_mean_abs) < 2e-2 * ref_mean_abs + 1e-3
    assert acts[1] is means and acts[2] is logstds
    # fused KL against the reference's value
    import bdvs_b200 as B
    kl = float(B.compute_kl_with_prior(means, logstds))
    assert abs(kl - fx["kl"]) < 5e-3 * abs(fx["kl"])


@pytest.mark.parametrize("name", ["tiny3", "tiny30"])
def test_test_forward_and_submodules_match_golden(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    x, c, eps, prior = O.make_inputs(cfg, cfg["batch"])
    with torch.no_grad():
        # prior sampling: inject the prior draws through torch.randn the way the module draws them
        import bdvs_b200.vunets as V
        draws = iter(cuda(prior))
        orig = V._draw_eps
        V._draw_eps = lambda shape, device: next(draws)
        try:
            pimgs = m.test_forward(c.cuda())
        finally:
            V._draw_eps = orig
        assert rel(pimgs, fx["test_forward"]) < BF16_TOL
        # vunet.ed(vunet.eu(x)) as experiments/shape_and_pose_net.py:413 calls it
        draws = iter(cuda(eps))
        V._draw_eps = lambda shape, device: next(draws)
        try:
            hs_in = m.eu(x.cuda())
            n_in = len(hs_in)
            hs, means, logstds, zs = m.ed(hs_in)
        finally:
            V._draw_eps = orig
    assert n_in - len(hs_in) == 5  # ed popped five of eu's outputs in place, like the reference
    assert [tuple(h.shape) for h in hs] == fx["ed_hs_shapes"]
    for i in range(2):
        assert rel(zs[i], fx["zs"][i]) < BF16_TOL
        assert rel(means[i], fx["means"][i]) < BF16_TOL


@pytest.mark.parametrize("cfg,batch", [(H256, 1), (D256, 2), (M128, 3)])
def test_full_size_configs_match_oracle(cfg, batch):
    """BASELINE.json configs at their real spatial sizes, small batch so the CPU oracle takes seconds."""
    m, sd = build(cfg)
    x, c, eps, _ = O.make_inputs(cfg, batch, tag="full")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        o_imgs, o_means, o_logstds, o_zs, _ = O.forward(sd, x, c, eps)
        o_t = O.transfer(sd, x, c, eps)
        imgs, means, logstds, _ = m(x.cuda(), c.cuda(), eps=cuda(eps))
        timgs = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps))
    assert rel(imgs, o_imgs) < BF16_TOL
    assert rel(timgs, o_t) < BF16_TOL
    for i in range(2):
        assert rel(means[i], o_means[i]) < BF16_TOL
        assert rel(logstds[i], o_logstds[i]) < BF16_TOL


def test_latent_noise_is_bit_exact_with_torch_generator():
    """The module draws eps with torch.randn in the reference's order/shapes: seeding the global generator
    gives bit-identical results to injecting those very draws (and therefore to the reference's
    randn_like stream on the same device)."""
    cfg = dict(H256, spatial_size=64)
    m, _ = build(cfg)
    x, c, _, _ = O.make_inputs(cfg, 4, tag="rng")
    x, c = x.cuda(), c.cuda()
    shp = O.latent_shapes(4, 64, O.n_scales_of(cfg)[0], 128)
    torch.manual_seed(1234)
    e0 = torch.randn(shp[0], device="cuda")
    e1 = torch.randn(shp[1], device="cuda")
    with torch.no_grad():
        a = m(x, c, eps=[e0, e1])
        torch.manual_seed(1234)
        b = m(x, c)
        torch.manual_seed(1234)
        t1 = m.transfer(x, c)
        t2 = m.transfer(x, c, eps=[e0, e1])
    assert torch.equal(a[0], b[0]) and torch.equal(a[1][1], b[1][1]) and torch.equal(a[2][0], b[2][0])
    assert torch.equal(t1, t2)
    # z = mu + exp(logstd) * eps exactly, with the injected eps
    with torch.no_grad():
        torch.manual_seed(1234)
        _, mu, ls, z = m.ed(m.eu(x))
    for i, e in enumerate((e0, e1)):
        assert rel(z[i], mu[i] + torch.exp(ls[i]) * e) < 1e-6


def test_batch_independence_and_determinism_at_full_size(monkeypatch):
    """Size-independent properties at H256 (BASELINE configs[1] shape): frames are independent units —
    a frame's result does not depend on what else is in the batch — and a rerun is bit-identical.  At B = 8 the
    128-channel layers run on the CTA-pair tile kernel, at B = 3 on the single-CTA / per-tap kernels: same bits."""
    m, _ = build(H256)
    B = 8
    x, c, eps, _ = O.make_inputs(H256, B, tag="prop")
    x, c, eps = x.cuda(), c.cuda(), cuda(eps)
    with torch.no_grad():
        full = m.transfer(x, c, eps=eps)
        again = m.transfer(x, c, eps=eps)
        part = m.transfer(x[2:5], c[2:5], eps=[e[2:5].contiguous() for e in eps])
    assert torch.equal(full, again)
    assert torch.equal(full[2:5], part)
    assert torch.isfinite(full).all()
    monkeypatch.setenv("VUB_PAIR", "0")  # single-CTA tile kernel for the layers the pair kernel took above
    with torch.no_grad():
        single = m.transfer(x, c, eps=eps)
    assert torch.equal(full, single)


def test_ops_reject_bad_inputs():
    import bdvs_b200 as B
    with pytest.raises(NotImplementedError):
        B.NormConv2d(48, 32, 3, padding=1).cuda()(torch.zeros(1, 48, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):  # non power-of-two extent is refused by the C ABI, not silently mis-tiled
        B.NormConv2d(32, 32, 3, padding=1).cuda()(torch.zeros(1, 32, 12, 12, device="cuda"))


def test_standalone_blocks_match_oracle():
    """Block-level parity through the reference-style public forward() of each module (NCHW in/out)."""
    import bdvs_b200 as B
    with contextlib.redirect_stdout(io.StringIO()):
        pass
    torch.manual_seed(0)
    x = O.det_uniform("blk_x", (2, 64, 16, 16))
    a = O.det_uniform("blk_a", (2, 64, 16, 16))
    # VunetRNB with skip
    blk = B.VunetRNB(channels=64, a_channels=64, residual=True)
    shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    sd = O.det_state_dict(shapes)
    blk.load_state_dict(sd)
    ref = O.rnb({("b." + k): v for k, v in sd.items()}, "b", x, a)
    got = blk.cuda()(x.cuda(), a.cuda())
    assert got.shape == ref.shape and rel(got, ref) < BF16_TOL
    # VunetRNB without skip, NormConv2d 3x3, Downsample, Upsample
    blk2 = B.VunetRNB(channels=64)
    sd2 = O.det_state_dict({k: tuple(v.shape) for k, v in blk2.state_dict().items()})
    blk2.load_state_dict(sd2)
    assert rel(blk2.cuda()(x.cuda()), O.rnb({("b." + k): v for k, v in sd2.items()}, "b", x)) < BF16_TOL
    nc = B.NormConv2d(64, 128, 3, padding=1)
    sd3 = O.det_state_dict({k: tuple(v.shape) for k, v in nc.state_dict().items()})
    nc.load_state_dict(sd3)
    assert rel(nc.cuda()(x.cuda()), O.norm_conv({("c." + k): v for k, v in sd3.items()}, "c", x)) < BF16_TOL
    dn = B.Downsample(64, 128)
    sd4 = O.det_state_dict({k: tuple(v.shape) for k, v in dn.state_dict().items()})
    dn.load_state_dict(sd4)
    assert rel(dn.cuda()(x.cuda()), O.norm_conv({("d." + k): v for k, v in sd4.items()}, "d.down", x, stride=2)) < BF16_TOL
    up = B.Upsample(64, 32)
    sd5 = O.det_state_dict({k: tuple(v.shape) for k, v in up.state_dict().items()})
    up.load_state_dict(sd5)
    assert rel(up.cuda()(x.cuda()), O.upsample({("u." + k): v for k, v in sd5.items()}, "u", x)) < BF16_TOL
    # stem-shaped 1x1 conv on an RGB image
    st = B.NormConv2d(3, 32, 1)
    sd6 = O.det_state_dict({k: tuple(v.shape) for k, v in st.state_dict().items()})
    st.load_state_dict(sd6)
    img = O.det_uniform("img", (2, 3, 32, 32))
    assert rel(st.cuda()(img.cuda()), O.norm_conv({("s." + k): v for k, v in sd6.items()}, "s", img)) < BF16_TOL


# ---------------------------------------------------------------------------------------------
FP32_TOL = 1e-4  # BASELINE.json north_star: "1e-4 in an fp32-accumulate validation mode"


@pytest.mark.parametrize("name", ["tiny3", "tiny30", "market128"])
def test_fp32_validation_mode_matches_reference_to_1e4(golden, name):
    import bdvs_b200 as B
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    x, c, eps, prior = O.make_inputs(cfg, cfg["batch"])
    with torch.no_grad(), B.fp32_validation(m):
        imgs, means, logstds, acts = m(x.cuda(), c.cuda(), eps=cuda(eps))
        timgs = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps))
    assert imgs.dtype == torch.float32 and acts[0][0].dtype == torch.float32
    errs = {"imgs": rel(imgs, fx["imgs"]), "transfer": rel(timgs, fx["transfer"])}
    for i in range(2):
        errs[f"mean{i}"] = rel(means[i], fx["means"][i])
        errs[f"logstd{i}"] = rel(logstds[i], fx["logstds"][i])
    kl = float(B.compute_kl_with_prior(means, logstds))
    errs["kl"] = abs(kl - fx["kl"]) / abs(fx["kl"])
    if "test_forward" in fx:
        import bdvs_b200.vunets as V
        draws = iter(cuda(prior))
        orig = V._draw_eps
        V._draw_eps = lambda shape, device: next(draws)  # the reference's randn_like draws, injected
        try:
            with torch.no_grad(), B.fp32_validation(m):
                pimgs = m.test_forward(c.cuda())
        finally:
            V._draw_eps = orig
        errs["test_forward"] = rel(pimgs, fx["test_forward"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < FP32_TOL, errs
    # leaving the mode restores the bf16 tensor-core path
    with torch.no_grad():
        imgs_bf16 = m(x.cuda(), c.cuda(), eps=cuda(eps))[0]
    assert 1e-4 < rel(imgs_bf16, fx["imgs"]) < BF16_TOL


@pytest.mark.parametrize("name", ["tiny3", "tiny30", "market128"])
def test_fp32_validation_mode_gradients_match_reference_to_1e4(golden, name):
    """north_star: images, latents AND gradients within 1e-4 "in an fp32-accumulate validation mode".  The backward
    schedule runs on the CUDA-core fp32 kernels (vub_conv_dgrad_f32 / vub_conv_wgrad_f32 / vub_latent_bwd_f32) and
    the same vub_wn_bwd as the bf16 path; checked against the CPU oracle's fp32 autograd (concatenated gradient and
    every single tensor) and against the gradient fixtures generated from the reference itself."""
    import bdvs_b200 as B
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    m.train()
    bs = cfg["batch"]
    x, c, eps, _ = O.make_inputs(cfg, bs)
    target = O.det_uniform("target", (bs, 3, cfg["spatial_size"], cfg["spatial_size"]))
    with B.fp32_validation(m):
        imgs, means, logstds, _ = m(x.cuda(), c.cuda(), eps=cuda(eps))
        assert imgs.requires_grad
        loss = torch.mean((imgs - target.cuda()) ** 2) + 1e-3 * B.compute_kl_with_prior(means, logstds)
        loss.backward()
    torch.cuda.synchronize()
    got = {k: (p.grad.detach().float().cpu() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_imgs, o_means, o_logstds, _, _ = O.forward(sd_g, x, c, eps)
    o_loss = torch.mean((o_imgs - target) ** 2) + 1e-3 * O.compute_kl_with_prior(o_means, o_logstds)
    o_loss.backward()
    ref = {k: v.grad for k, v in sd_g.items()}
    assert abs(float(loss) - float(o_loss)) < 1e-4 * abs(float(o_loss))
    assert abs(float(loss) - fx["loss"]) < 1e-4 * abs(fx["loss"])
    assert sorted(k for k, g in got.items() if g is None) == fx["grad_is_none"]
    keys = [k for k, g in ref.items() if g is not None and float(g.abs().max()) > 0.0]
    num = sum(float((got[k] - ref[k]).pow(2).sum()) for k in keys)
    den = sum(float(ref[k].pow(2).sum()) for k in keys)
    total = (num / den) ** 0.5
    worst = max(float((got[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-20)) for k in keys)
    fixture = max(float((got[k] - g).norm() / g.norm()) for k, g in fx["grad_full"].items())
    print(f"{name}: fp32 validation gradients: concatenated rel-L2 {total:.2e}, worst tensor {worst:.2e}, "
          f"vs reference fixture tensors {fixture:.2e}")
    assert total < FP32_TOL
    assert worst < 20 * FP32_TOL       # tiny tensors (a bias of 32 values) carry proportionally more round-off
    assert fixture < FP32_TOL * 10
    for k, n_ref in fx["grad_norms"].items():
        assert abs(float(got[k].norm()) - n_ref) < 1e-3 * n_ref + 1e-9, k
    # leaving the mode: the bf16 tensor-core backward is back (same buffers are not reused across precisions)
    m.zero_grad(set_to_none=True)
    imgs2 = m(x.cuda(), c.cuda(), eps=cuda(eps))[0]
    ((imgs2 - target.cuda()) ** 2).mean().backward()
    g2 = m.dd.out_conv.conv.weight_v.grad.float().cpu()
    assert 1e-5 < float((g2 - ref["dd.out_conv.conv.weight_v"]).norm() / ref["dd.out_conv.conv.weight_v"].norm()) < 5e-2


def test_sequence_level_encoder_reuse_equals_per_frame_transfer():
    """encode_appearance once per sequence + decode_pose over its frames == transfer on the repeated appearance
    image with the same latent noise."""
    from bdvs_b200.synthesis import SequenceSynthesizer
    cfg = dict(H256, spatial_size=64)
    m, _ = build(cfg)
    S, T = 2, 3
    x, _, eps, _ = O.make_inputs(cfg, S, tag="seq")
    _, c, _, _ = O.make_inputs(cfg, S * T, tag="seqc")
    with torch.no_grad():
        app = m.encode_appearance(x.cuda(), eps=cuda(eps))
        seq = m.decode_pose(c.cuda(), app, T)
        ref = m.transfer(x.cuda().repeat_interleave(T, 0), c.cuda(), eps=[e.cuda().repeat_interleave(T, 0) for e in eps])
    assert torch.equal(seq, ref)
    synth = SequenceSynthesizer(m, S, T)
    out = synth(x.cuda(), c.cuda())
    assert out.shape == ref.shape and bool(torch.isfinite(out).all())
    with pytest.raises(ValueError):
        m.decode_pose(c.cuda()[:4], app, T)


def test_uint8_frame_epilogue_is_bit_exact_with_reference_expression():
    """(scale_img(rgb) * 255).permute(0, 2, 3, 1).astype(uint8) — lib/utils.py:658-668,
    data/data_conversions_3d.py:1166-1173."""
    from bdvs_b200 import ops
    from bdvs_b200.synthesis import FrameSynthesizer
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(3, 3, 32, 64, generator=g) * 3 - 1.5)  # includes values outside [-1, 1]
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.9999999])
    ref = (torch.clamp((x + 1.0) / 2.0, 0.0, 1.0) * 255.0).permute(0, 2, 3, 1).numpy().astype("uint8")
    got = ops.frames_to_uint8(x.cuda()).cpu().numpy()
    assert got.shape == ref.shape and (got == ref).all()
    cfg = dict(H256, spatial_size=64)
    m, _ = build(cfg)
    fs = FrameSynthesizer(m, 2, uint8_frames=True)
    out = fs.run()
    assert out.dtype == torch.uint8 and tuple(out.shape) == (2, 64, 64, 3)


@pytest.mark.parametrize("cfg,batch", [(H256, 1), (M128, 2)])
def test_activation_layouts_and_kernel_variants_agree(cfg, batch, monkeypatch):
    """The inference path has several equivalent executions: lean activations (ELU on load) vs (raw, ELU) pairs,
    A operand in tensor vs shared memory, one stream vs two, CTA-pair vs single-CTA tile kernel.  All must sit inside the bf16 tolerance of the oracle and
    close to each other (they differ by one bf16 rounding of the ELU input at most)."""
    m, sd = build(cfg)
    x, c, eps, _ = O.make_inputs(cfg, batch, tag="variants")
    with torch.no_grad():
        o_t = O.transfer(sd, x, c, eps)
    outs = {}
    for name, env in (("default", {}), ("pairs", {"VUB_LEAN_ACTS": "0"}), ("smem_operand", {"VUB_ROWS_ATM": "0"}),
                      ("one_stream", {"VUB_FWD_STREAMS": "1"}), ("single_cta", {"VUB_PAIR": "0"})):
        for k in ("VUB_LEAN_ACTS", "VUB_ROWS_ATM", "VUB_FWD_STREAMS", "VUB_PAIR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with torch.no_grad():
            outs[name] = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps)).float().cpu()
        assert rel(outs[name], o_t) < BF16_TOL, name
    for name in ("pairs", "smem_operand", "one_stream", "single_cta"):
        assert rel(outs[name], outs["default"]) < 5e-3, name
    assert torch.equal(outs["one_stream"], outs["default"])          # same kernels, different stream assignment
    assert torch.equal(outs["smem_operand"], outs["default"])        # same operands, fetched differently
    assert torch.equal(outs["single_cta"], outs["default"])          # same K order, M = 128 vs M = 256 instructions


def test_uint8_image_inputs_are_bit_identical_to_converted_fp32_inputs():
    """uint8 HWC images through the stem kernels == the reference's ``ToTensor()`` + ``x * 2 - 1`` conversion
    (data/__init__.py:15-43) followed by the fp32 NCHW path, bit for bit — for the 3-channel stickman / appearance
    (H3.6M) and the 30-channel part crops (DeepFashion / Market)."""
    for cfg in (dict(H256, spatial_size=64), dict(D256, spatial_size=128)):
        m, _ = build(cfg)
        B = 3
        g = torch.Generator().manual_seed(11)
        s = cfg["spatial_size"]
        sx = s // 2 ** (m.n_scales - m.n_scales_x)
        x8 = torch.randint(0, 256, (B, sx, sx, cfg["n_channels_x"]), dtype=torch.uint8, generator=g)
        c8 = torch.randint(0, 256, (B, s, s, 3), dtype=torch.uint8, generator=g)
        x8[0, 0, 0, :3] = torch.tensor([0, 255, 128], dtype=torch.uint8)
        to_f = lambda t: ((t.float().div(255) * 2.0) - 1.0).permute(0, 3, 1, 2).contiguous()  # noqa: E731
        _, _, eps, _ = O.make_inputs(cfg, B, tag="u8")
        with torch.no_grad():
            a = m.transfer(x8.cuda(), c8.cuda(), eps=cuda(eps))
            b = m.transfer(to_f(x8).cuda(), to_f(c8).cuda(), eps=cuda(eps))
            f8 = m(x8.cuda(), c8.cuda(), eps=cuda(eps))
            f32 = m(to_f(x8).cuda(), to_f(c8).cuda(), eps=cuda(eps))
        assert torch.equal(a, b)
        assert torch.equal(f8[0], f32[0]) and torch.equal(f8[1][1], f32[1][1])


def test_host_pipeline_uint8_matches_direct_transfer():
    """``FrameSynthesizer.render_host`` with uint8 HWC host buffers (two alternating graph / buffer sets, copies on
    their own streams, no staging copies) returns exactly ``transfer(x, c, uint8_frames=True)`` for every batch of a
    stream of different batches.  The latent noise is drawn inside the step from the global generator: an eager
    synthesizer must reproduce the eager loop bit for bit, and the graphed pipeline must reproduce the same
    synthesizer's sequential graph replays (same seed -> same Philox offsets per replay)."""
    from bdvs_b200.synthesis import FrameSynthesizer
    cfg = dict(H256, spatial_size=64)
    m, _ = build(cfg)
    B, n_batches = 4, 5
    g = torch.Generator().manual_seed(3)
    xs = [torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_batches)]
    cs = [torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_batches)]

    def pipeline(fs):
        outs = [torch.zeros(B, 64, 64, 3, dtype=torch.uint8).pin_memory() for _ in range(n_batches)]
        torch.manual_seed(99)
        for i in range(n_batches):
            fs.render_host(xs[i], cs[i], outs[i])
        fs.join()
        torch.cuda.synchronize()
        return outs

    # (a) eager steps through the pipeline == the plain eager loop
    fs_eager = FrameSynthesizer(m, B, uint8_frames=True, uint8_inputs=True, use_graph=False)
    pipeline(fs_eager)  # the first call sets the second buffer set up (its warm-up steps draw noise): settle first
    outs = pipeline(fs_eager)
    torch.manual_seed(99)
    with torch.no_grad():
        for i in range(n_batches):
            want = m.transfer(xs[i].cuda(), cs[i].cuda(), uint8_frames=True).cpu()
            assert torch.equal(outs[i], want), i
    # (b) graphed pipeline (two alternating sets) == sequential replays of the same synthesizer
    fs = FrameSynthesizer(m, B, uint8_frames=True, uint8_inputs=True)
    pipeline(fs)  # the first call captures the second graph (its warm-up steps draw noise): settle before seeding
    outs = pipeline(fs)
    torch.manual_seed(99)
    for i in range(n_batches):
        want = fs(xs[i].cuda(), cs[i].cuda()).cpu()
        assert torch.equal(outs[i], want), i
    assert len({int(o.sum()) for o in outs}) == n_batches  # different batches really produced different frames
    with pytest.raises(ValueError):
        fs.render_host(xs[0].float(), cs[0], outs[0])
    # fp32 NCHW pipeline
    fs32 = FrameSynthesizer(m, B, use_graph=False)
    x, c, _, _ = O.make_inputs(cfg, B, tag="host32")
    out32 = torch.zeros(B, 3, 64, 64).pin_memory()
    fs32.render_host(x.pin_memory(), c.pin_memory(), out32)  # settle (see above)
    fs32.join()
    torch.manual_seed(5)
    fs32.render_host(x.pin_memory(), c.pin_memory(), out32)
    fs32.join()
    torch.cuda.synchronize()
    torch.manual_seed(5)
    with torch.no_grad():
        assert torch.equal(out32, m.transfer(x.cuda(), c.cuda()).cpu())


@pytest.mark.parametrize("cfg", [H256, M128])
def test_out_conv_uint8_epilogue_equals_converted_fp32_image(cfg):
    """``transfer(..., uint8_frames=True)``: at >= 128 pixels wide ``out_conv``'s epilogue stores the uint8 HWC video
    frame itself (VUB_CONV_U8_HWC); the bytes must equal the reference expression applied to the fp32 image."""
    from bdvs_b200 import ops
    m, _ = build(cfg)
    x, c, eps, _ = O.make_inputs(cfg, 2, tag="u8out")
    assert ops.u8_out_ok(1, 32, 16, 3, 1, cfg["spatial_size"], cfg["spatial_size"])
    with torch.no_grad():
        img = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps))
        frames = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps), uint8_frames=True)
    want = (torch.clamp((img + 1.0) / 2.0, 0.0, 1.0) * 255.0).permute(0, 2, 3, 1).cpu().numpy().astype("uint8")
    assert frames.dtype == torch.uint8 and tuple(frames.shape) == (2, cfg["spatial_size"], cfg["spatial_size"], 3)
    assert (frames.cpu().numpy() == want).all()
