"""GPU (-m gpu): behavior-net sampling (BASELINE.json configs[3]) on the CUDA path — the flow's reverse pass and the
residual LSTM decoder, every matrix product a tcgen05 "linear layer" launch — against the golden vector generated from
the unmodified reference classes and against the CPU oracle at the real sizes of config/behavior_net.yaml
(dim_hidden_b 1024, 15 flow blocks with 2048-wide MLPs, 50-step sequences, batch 256).  Tolerance: rel-L2 <= 1e-2 (bf16)."""
import pytest
import torch

import behavior_oracle as BO
import vunet_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make(cfg):
    import bdvs_b200 as B
    flow = B.UnsupervisedTransformer2(flow_in_channels=cfg["dim_hidden_b"], n_flows=cfg["n_flows"],
                                      flow_hidden_depth=cfg["flow_hidden_depth"],
                                      flow_mid_channels=cfg["dim_hidden_b"] * cfg["flow_mid_channels_factor"])
    net = B.ResidualBehaviorNet(n_kps=cfg["n_kps"], dim_hidden_b=cfg["dim_hidden_b"], decoder_arch="lstm",
                                linear_in_decoder=False)
    fsd = BO.det_flow_state_dict({k: tuple(v.shape) for k, v in flow.state_dict().items()})
    nsd = BO.det_net_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()})
    flow.load_state_dict(fsd)
    net.load_state_dict(nsd)
    return flow.cuda().eval(), net.cuda().eval(), fsd, nsd


def test_matches_reference_golden(golden):
    import bdvs_b200 as B
    fx = golden("behavior_tiny")
    cfg = fx["config"]
    flow, net, fsd, nsd = make(cfg)
    Bs, T, n, D = cfg["batch"], cfg["length"], cfg["n_kps"], cfg["dim_hidden_b"]
    z = O.det_normal("bz", (Bs, D)).cuda()
    x_pose = O.det_uniform("bx", (Bs, T, n)).cuda()
    with torch.no_grad():
        b = flow.reverse(z)
        assert tuple(b.shape) == (Bs, D, 1, 1)
        xs, cs, zs, bb = net.generate_seq(b.squeeze(-1).squeeze(-1), x_pose, len=T, start_frame=0)
    errs = {"b": rel(b.reshape(Bs, D), fx["b"]), "xs": rel(xs, fx["xs"]), "cs": rel(cs, fx["cs"])}
    print("behavior_tiny:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    assert zs == [] and tuple(xs.shape) == (Bs, T, n)
    # behavior inference from a sequence (the LSTM encoder) and the net's forward = infer_b + generate_seq
    seq = O.det_uniform("bseq", (Bs, T + 3, n)).cuda()
    with torch.no_grad():
        b_inf = net.infer_b(seq, False)
        fx_xs, fx_cs, _, fx_b = net(seq, x_pose, T, start_frame=0)
        o_xs, _ = BO.generate_seq(nsd, fx["infer_b"], x_pose[:, 0].cpu(), T)
    e_inf, e_fwd = rel(b_inf, fx["infer_b"]), rel(fx_xs, o_xs)
    print(f"  infer_b {e_inf:.2e}, forward (behavior transfer) {e_fwd:.2e}")
    assert e_inf < TOL and e_fwd < TOL and torch.equal(fx_b, b_inf)
    # the graph-captured sampler equals the two calls; flow(z, reverse=True) and sample() are the same path
    smp = B.PoseSequenceSampler(flow, net, Bs, T)
    out = smp(z, x_pose[:, :1])
    assert torch.equal(out, xs)
    assert torch.equal(flow(z, reverse=True), b)
    torch.manual_seed(5)
    s1 = flow.sample((Bs, D), device="cuda")
    torch.manual_seed(5)
    s2 = flow.reverse(torch.randn((Bs, D)).cuda()).reshape(Bs, D)
    assert torch.equal(s1, s2)


def test_sampler_graph_follows_weight_updates(golden):
    """A captured PoseSequenceSampler must not replay stale weights: the reference re-reads its parameters in every
    call.  In-place updates (optimiser step, load_state_dict) re-prepare into the SAME packed buffers — no re-capture."""
    import bdvs_b200 as B
    cfg = golden("behavior_tiny")["config"]
    flow, net, fsd, nsd = make(cfg)
    Bs, T, n, D = cfg["batch"], cfg["length"], cfg["n_kps"], cfg["dim_hidden_b"]
    z = O.det_normal("bz", (Bs, D)).cuda()
    x0 = O.det_uniform("bx", (Bs, 1, n)).cuda()
    smp = B.PoseSequenceSampler(flow, net, Bs, T)
    before = smp(z, x0).clone()
    graph, bank = smp.graph, flow._prep.bank
    with torch.no_grad():
        for p in list(flow.parameters())[:40]:
            p.mul_(1.05)
        net.decoder.n_out.weight.mul_(0.9)
        net.decoder.rnn.bias_hh.add_(0.01)
    after = smp(z, x0).clone()
    assert smp.graph is graph and flow._prep.bank is bank          # same graph, same packed buffers
    with torch.no_grad():
        fresh = net.generate_seq(flow.reverse(z).reshape(Bs, D), x0, T, 0)[0]
        sd_f = {k: v.detach().cpu() for k, v in flow.state_dict().items()}
        sd_n = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        o_xs, _ = BO.generate_seq(sd_n, BO.flow_reverse(sd_f, z.cpu(), cfg["n_flows"], cfg["flow_hidden_depth"]),
                                  x0[:, 0].cpu(), T)
    assert torch.equal(after, fresh) and not torch.equal(after, before)
    assert rel(after, o_xs) < TOL
    # writes autograd cannot see (p.data): bump_weight_epoch() is the documented call
    net.decoder.n_out.bias.data.add_(0.05)
    B.bump_weight_epoch()
    moved = smp(z, x0)
    assert not torch.equal(moved, after)
    with torch.no_grad():
        assert torch.equal(moved, net.generate_seq(flow.reverse(z).reshape(Bs, D), x0, T, 0)[0])


def test_full_size_matches_oracle():
    """config/behavior_net.yaml: dim_hidden_b 1024, 15 blocks, MLP width 2048, depth 2; batch 256, 50 steps."""
    cfg = dict(dim_hidden_b=1024, flow_mid_channels_factor=2, n_flows=15, flow_hidden_depth=2, n_kps=51)
    flow, net, fsd, nsd = make(cfg)
    Bs, T = 256, 50
    z = O.det_normal("bz_full", (Bs, 1024))
    x0 = O.det_uniform("bx_full", (Bs, 1, 51))
    with torch.no_grad():
        b = flow.reverse(z.cuda()).reshape(Bs, 1024)
        xs, cs, _, _ = net.generate_seq(b, x0.cuda(), T, 0)
        o_b = BO.flow_reverse(fsd, z, 15, 2)
        o_xs, o_cs = BO.generate_seq(nsd, o_b, x0[:, 0], T)
        # the decoder alone from the oracle's b (isolates it from the flow's rounding)
        xs2, _, _, _ = net.generate_seq(o_b.cuda(), x0.cuda(), T, 0)
        seq = O.det_uniform("bseq_full", (Bs, T, 51))
        b_inf = net.infer_b(seq.cuda(), False)
        o_inf = BO.infer_b(nsd, seq)
    errs = {"b": rel(b, o_b), "xs": rel(xs, o_xs), "cs": rel(cs, o_cs), "xs_from_oracle_b": rel(xs2, o_xs),
            "velocities": rel(xs - cs, o_xs - o_cs), "infer_b (50-step LSTM-1024 encoder)": rel(b_inf, o_inf)}
    print("behavior net, full size B=256 T=50:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    assert torch.isfinite(xs).all()


def test_behavior_video_renderer_equals_the_separate_calls(golden):
    """configs[4] in one object: sampled pose sequences -> projection -> stickmen -> sequence synthesis == the same steps
    issued by hand (PoseSequenceSampler, project_poses, StickmanRasterizer, encode_appearance + decode_pose)."""
    import collections
    import contextlib
    import io
    import bdvs_b200 as B
    fx = golden("behavior_tiny")
    cfg = dict(fx["config"])
    flow, net, fsd, nsd = make(cfg)
    vfx = golden("tiny3")
    vcfg = {k: v for k, v in vfx["config"].items() if k != "batch"}
    with contextlib.redirect_stdout(io.StringIO()):
        vunet = B.VunetAlter(**vcfg)
    vunet.load_state_dict(O.det_state_dict({k: tuple(v.shape) for k, v in vunet.state_dict().items()}))
    vunet = vunet.cuda().eval()
    JM = collections.namedtuple("JointModel", "body right_lines left_lines head_lines face rshoulder lshoulder headup")
    jm = JM(body=[0, 14, 8, 11, 3], right_lines=[(0, 1), (1, 2), (0, 14), (14, 15), (15, 16)],
            left_lines=[(3, 4), (4, 5), (3, 11), (11, 12), (12, 13)], head_lines=[(8, 9), (9, 10)], face=[],
            rshoulder=14, lshoulder=11, headup=10)
    extr = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 4.0]], dtype=torch.float64)
    intr = (60.0, 16.0, 60.0, 16.0)
    T, S = 4, 4
    r = B.BehaviorVideoRenderer(vunet, flow, net, jm, extr, intr, (32.0, 32.0), frames=T, n_sampled=S)
    z = O.det_normal("bz", (S, cfg["dim_hidden_b"])).cuda()
    x0 = O.det_uniform("bx", (S, 1, cfg["n_kps"])).cuda()
    app = (O.det_uniform("app", (S, 32, 32, 3), 0.0, 255.0)).to(torch.uint8).cuda()
    torch.manual_seed(11)
    joints = r.sample_poses(z, x0)
    assert tuple(joints.shape) == (S, T, 17, 2)
    frames = r.render_sequence(app[1:2], joints[1]).clone()
    # by hand
    with torch.no_grad():
        b = flow.reverse(z).reshape(S, -1)
        poses = net.generate_seq(b, x0, T, 0)[0]
        j2 = B.project_poses(poses.reshape(S * T, 17, 3), extr, intr, (32.0, 32.0), 32).reshape(S, T, 17, 2)
        assert torch.equal(j2, joints)
        stick = B.StickmanRasterizer(jm, 32)(j2[1])
        torch.manual_seed(11)
        ref = vunet.decode_pose(stick, vunet.encode_appearance(app[1:2]), T, uint8_frames=True)
    assert frames.dtype == torch.uint8 and tuple(frames.shape) == (T, 32, 32, 3)
    # the synthesizer draws its latent noise inside the graph: same distribution, not the same bits as the eager call
    diff = (frames.float() - ref.float()).abs().mean() / 255.0
    assert float(diff) < 0.1
    assert int(stick.max()) > 0   # the sampled poses land inside the image
