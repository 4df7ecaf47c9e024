"""End-to-end flow parity: this package's RAFT (cuDNN convs + sm_100a correlation kernels) against
flows produced by the REFERENCE RAFT on the same name-seeded weights and seeded frames
(tests/golden/raft.npz).  Tolerance: EPE <= 1e-2 px mean for the fp32-faithful and
tf32 correlation modes; convolutions run in true fp32 here so the comparison isolates the path."""
import numpy as np
import pytest
import torch

from tests import golden_inputs as gi

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _engine(name, cuda, **kw):
    from sd_animation_optical_flow_b200.engine import RaftEngine
    cfg = gi.RAFT_CASES[name]
    return RaftEngine(checkpoint=None, iters=cfg['iters'], small=cfg['small'], seed=cfg['seed'], device=cuda, **kw)


def _epe(a, b):
    return np.sqrt(((a - b) ** 2).sum(0))


@pytest.mark.parametrize('name', list(gi.RAFT_CASES))
@pytest.mark.parametrize('mode', ['fp32', '3xtf32', 'tf32', 'fp16', 'alt'])
def test_flow_matches_reference_raft(cuda, golden, name, mode):
    kw = dict(alternate_corr=True) if mode == 'alt' else dict(corr_precision=mode)
    eng = _engine(name, cuda, **kw)
    img1, img2 = gi.raft_inputs(name)
    a = torch.from_numpy(img1).to(cuda)[None]
    b = torch.from_numpy(img2).to(cuda)[None]
    flow = eng.estimate_flow(a, b, unpad=False)[0].permute(2, 0, 1).cpu().numpy()
    ref = golden['raft'][f'{name}_flow_up']
    assert flow.shape == ref.shape
    epe = _epe(flow, ref)
    print(f'{name}/{mode}: EPE mean {epe.mean():.2e} max {epe.max():.2e} (mean |flow| {np.abs(ref).mean():.1f} px)')
    assert epe.mean() <= 1e-2 and epe.max() <= 1e-1


def test_fast_nhwc_forward_equals_module_forward(cuda, golden):
    """raft_fast.FastRaft (hand-scheduled NHWC loop + glue kernels) against the plain nn.Module forward with the
    same weights and correlation mode, and against the reference's flow."""
    cfg = gi.RAFT_CASES['basic']
    fast = _engine('basic', cuda, corr_precision='3xtf32', fast=True)
    slow = _engine('basic', cuda, corr_precision='3xtf32', fast=False)
    assert fast.fast is not None and slow.fast is None
    img1, img2 = gi.raft_inputs('basic')
    a = torch.from_numpy(img1).to(cuda)[None]
    b = torch.from_numpy(img2).to(cuda)[None]
    f_fast = fast.estimate_flow(a, b, unpad=False)
    f_slow = slow.estimate_flow(a, b, unpad=False)
    d = (f_fast - f_slow).norm(dim=-1)
    print(f'fast vs module: EPE mean {float(d.mean()):.2e} max {float(d.max()):.2e}')
    assert float(d.mean()) <= 2e-3 and float(d.max()) <= 2e-2
    epe = _epe(f_fast[0].permute(2, 0, 1).cpu().numpy(), golden['raft']['basic_flow_up'])
    assert epe.mean() <= 1e-2 and epe.max() <= 1e-1
    # two pairs at once
    both = fast.estimate_flow(torch.cat([a, b]), torch.cat([b, a]), unpad=False)
    assert float((both[:1] - f_fast).abs().max()) <= 2e-3


def test_instnorm_relu_kernel(cuda):
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(4)
    for shape in ((2, 8, 37, 41), (1, 64, 96, 64), (2, 3, 5, 3)):
        x = torch.randn(shape, generator=g, device=cuda) * 3 + 1.5
        ref = torch.nn.functional.instance_norm(x, eps=1e-5)
        assert torch.allclose(ops.instnorm_relu(x.clone(), relu=False), ref, atol=2e-5, rtol=1e-5)
        assert torch.allclose(ops.instnorm_relu(x, relu=True, inplace=False), torch.relu(ref), atol=2e-5, rtol=1e-5)


def test_instnorm_nhwc_kernels(cuda):
    """channels-last InstanceNorm (+ReLU, + residual tail) vs F.instance_norm; tolerance 2e-5 abs (fp32 summation order)."""
    from sd_animation_optical_flow_b200 import ops
    CL = torch.channels_last
    g = torch.Generator(device=cuda).manual_seed(5)
    for shape in ((2, 64, 37, 41), (1, 96, 48, 32), (2, 128, 12, 8), (3, 8, 5, 3)):
        N, C = shape[:2]
        x = (torch.randn(shape, generator=g, device=cuda) * 3 + 1.5).contiguous(memory_format=CL)
        res = torch.randn(shape, generator=g, device=cuda).contiguous(memory_format=CL)
        ref = torch.nn.functional.instance_norm(x, eps=1e-5)
        stats = lambda: torch.zeros((N * C * 2,), dtype=torch.float64, device=cuda)
        y = ops.instnorm_nhwc(x.clone(memory_format=torch.preserve_format), stats(), relu=False)
        assert y.is_contiguous(memory_format=CL) and torch.allclose(y, ref, atol=2e-5, rtol=1e-5)
        y = ops.instnorm_nhwc(x.clone(memory_format=torch.preserve_format), stats(), relu=True)
        assert torch.allclose(y, torch.relu(ref), atol=2e-5, rtol=1e-5)
        y = ops.instnorm_nhwc(x.clone(memory_format=torch.preserve_format), stats(), relu=True, residual=res)
        assert torch.allclose(y, torch.relu(res + torch.relu(ref)), atol=2e-5, rtol=1e-5)
        z = ops.add_relu_(x.clone(memory_format=torch.preserve_format), res)
        assert torch.equal(z, torch.relu(x + res))
    with pytest.raises(RuntimeError):
        ops.instnorm_nhwc(torch.randn((1, 8, 4, 4), device=cuda), torch.zeros(16, dtype=torch.float64, device=cuda))  # NCHW


def test_small_conv_kernels(cuda):
    """conv7x7_c2_relu (convf1) and flowhead2_update (flow_head.conv2 + coords update) vs F.conv2d in fp32.
    Tolerance 2e-5 abs relative to unit-scale activations (fp32 FMA, different summation order)."""
    import torch.nn.functional as F
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(6)
    for (B, h, w) in ((1, 96, 64), (2, 13, 21), (1, 8, 8)):
        flow = torch.randn((B, h, w, 2), generator=g, device=cuda) * 4
        wt = torch.randn((128, 2, 7, 7), generator=g, device=cuda) * 0.1
        bias = torch.randn((128,), generator=g, device=cuda)
        ref = F.relu(F.conv2d(flow.permute(0, 3, 1, 2), wt, bias, padding=3)).permute(0, 2, 3, 1)
        out = ops.conv7x7_c2_relu(flow, wt.permute(2, 3, 1, 0).contiguous(), bias)
        assert out.shape == ref.shape
        assert float((out - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
        x = torch.relu(torch.randn((B, h, w, 256), generator=g, device=cuda))
        w2 = torch.randn((2, 256, 3, 3), generator=g, device=cuda) * 0.05
        b2 = (0.3, -0.7)
        delta = F.conv2d(x.permute(0, 3, 1, 2), w2, torch.tensor(b2, device=cuda), padding=1).permute(0, 2, 3, 1)
        ys, xs = torch.meshgrid(torch.arange(h, device=cuda), torch.arange(w, device=cuda), indexing='ij')
        grid = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1)
        coords1 = (grid + torch.randn((B, h, w, 2), generator=g, device=cuda)).contiguous()
        want_c = coords1 + delta
        fl = torch.empty((B, h, w, 2), device=cuda)
        hx = torch.zeros((B, h, w, 12), device=cuda)
        rhx = torch.zeros((B, h, w, 12), device=cuda)
        ops.flowhead2_update(x, w2.permute(2, 3, 0, 1).contiguous(), b2, coords1, fl, hx, 10, rhx, 6)
        assert float((coords1 - want_c).abs().max()) <= 5e-5
        assert float((fl - (want_c - grid)).abs().max()) <= 5e-5
        assert torch.equal(hx[..., 10:12], fl) and torch.equal(rhx[..., 6:8], fl) and float(hx[..., :10].abs().max()) == 0


def test_convex_upsample_kernel(cuda):
    from sd_animation_optical_flow_b200 import ops
    from sd_animation_optical_flow_b200.raft import convex_upsample
    g = torch.Generator(device=cuda).manual_seed(3)
    flow = torch.randn((2, 2, 12, 20), generator=g, device=cuda) * 3
    mask = torch.randn((2, 576, 12, 20), generator=g, device=cuda) * 2
    ref = convex_upsample(flow, 0.25 * mask)
    out = ops.convex_upsample(mask.permute(0, 2, 3, 1).contiguous(), flow.permute(0, 2, 3, 1).contiguous(), 0.25)
    assert torch.allclose(out.permute(0, 3, 1, 2), ref, atol=2e-5, rtol=1e-5)


def test_bf16_volume_flow_stays_close(cuda, golden):
    eng = _engine('basic', cuda, corr_precision='bf16')
    img1, img2 = gi.raft_inputs('basic')
    flow = eng.estimate_flow(torch.from_numpy(img1).to(cuda)[None], torch.from_numpy(img2).to(cuda)[None], unpad=False)
    epe = _epe(flow[0].permute(2, 0, 1).cpu().numpy(), golden['raft']['basic_flow_up'])
    print(f'bf16: EPE mean {epe.mean():.2e} max {epe.max():.2e}')
    assert epe.mean() <= 5e-2


def test_raft2_calc_dropin(cuda, golden):
    """RAFT_2.calc(img1_bgr, img2_bgr) -> float32 [H',W',2] of the PADDED size (ofgen.py:70-79)."""
    from sd_animation_optical_flow_b200 import ofgen
    cfg = gi.RAFT_CASES['basic_pad']
    algo = ofgen.RAFT_2(model_path=None, iters=cfg['iters'], seed=cfg['seed'], corr_precision='3xtf32')
    img1, img2 = gi.raft_inputs('basic_pad')
    flow, v = ofgen.of_calc(img1[:, :, ::-1], img2[:, :, ::-1], algo)      # scripts pass BGR
    ref = golden['raft']['basic_pad_flow_up'].transpose(1, 2, 0)
    assert flow.shape == ref.shape == (136, 152, 2) and flow.dtype == np.float32
    assert _epe(flow.transpose(2, 0, 1), ref.transpose(2, 0, 1)).mean() <= 1e-2
    np.testing.assert_allclose(v, np.sqrt((flow ** 2).sum(-1)), rtol=1e-6)
    with pytest.raises(FileNotFoundError):
        ofgen.RAFT_2('RAFT/models/raft-things.pth')


def test_cuda_graph_replay_equals_eager(cuda):
    eager = _engine('basic', cuda, use_cuda_graph=False)
    graphed = _engine('basic', cuda, use_cuda_graph=True)
    img1, img2 = gi.raft_inputs('basic')
    a = torch.from_numpy(img1).to(cuda)[None]
    b = torch.from_numpy(img2).to(cuda)[None]
    f0 = eager.estimate_flow(a, b)
    f1 = graphed.estimate_flow(a, b)
    f2 = graphed.estimate_flow(b, a)      # replay with new inputs
    f3 = graphed.estimate_flow(a, b)
    # replays are bit-identical except for the order of the fp64 atomics behind the instance-norm statistics
    # (1e-16 relative, below fp32 resolution except on a rounding boundary)
    assert torch.allclose(f0, f1, atol=1e-4) and torch.allclose(f1, f3, atol=1e-5) and not torch.allclose(f1, f2, atol=1e-2)


def test_batched_pairs_equal_single_pairs(cuda):
    eng = _engine('basic', cuda, corr_precision='3xtf32')
    i1, i2 = gi.raft_inputs('basic')
    j1, j2 = gi.shifted_pair(128, 160, 555, dx=-2, dy=5)
    a = torch.from_numpy(np.stack([i1, j1])).to(cuda)
    b = torch.from_numpy(np.stack([i2, j2])).to(cuda)
    both = eng.estimate_flow(a, b)
    one = eng.estimate_flow(a[1:], b[1:])
    assert float((both[1:] - one).abs().max()) <= 2e-3


def test_config5_size_fast_path_equals_module_forward(cuda):
    """720x1280 (90x160 features, N = 14400: partial correlation tiles, odd pooled sizes 45x80 / 22x40 / 11x20) through
    the uint8 fast path with a CUDA graph vs the plain module forward; tolerance as above (EPE <= 2e-3 mean)."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    f1, f2 = gi.shifted_pair(720, 1280, 77, dx=5, dy=-2)
    a = torch.from_numpy(f1).to(cuda)[None]
    b = torch.from_numpy(f2).to(cuda)[None]
    fast = RaftEngine(checkpoint=None, iters=6, seed=1, device=cuda, corr_precision='3xtf32', use_cuda_graph=True)
    slow = RaftEngine(checkpoint=None, iters=6, seed=1, device=cuda, corr_precision='3xtf32', fast=False)
    ff = fast.estimate_flow(a, b)
    fs = slow.estimate_flow(a, b)
    assert ff.shape == fs.shape == (1, 720, 1280, 2)
    d = (ff - fs).norm(dim=-1)
    print(f'720x1280 fast vs module: EPE mean {float(d.mean()):.2e} max {float(d.max()):.2e}')
    assert float(d.mean()) <= 2e-3 and float(d.max()) <= 5e-2
    assert torch.allclose(fast.estimate_flow(a, b), ff, atol=1e-5)     # graph replay
    # a frame size that needs padding (InputPadder) through the fused normalise+pad kernel
    g1, g2 = gi.shifted_pair(250, 333, 78)
    c, d2 = torch.from_numpy(g1).to(cuda)[None], torch.from_numpy(g2).to(cuda)[None]
    e = (fast.estimate_flow(c, d2) - slow.estimate_flow(c, d2)).norm(dim=-1)
    assert fast.estimate_flow(c, d2).shape == (1, 250, 333, 2) and float(e.mean()) <= 2e-3
    assert fast.estimate_flow(c, d2, unpad=False).shape == (1, 256, 336, 2)


def test_graph_output_is_a_copy_when_only_rows_are_padded(cuda):
    """B=1 and height-only padding (132x160 -> 136x160): the unpadded slice of the graph's static output
    is 'contiguous', so it must be cloned explicitly -- a second call must not overwrite the first result."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    f1, f2 = gi.shifted_pair(132, 160, 91)
    a = torch.from_numpy(f1).to(cuda)[None]
    b = torch.from_numpy(f2).to(cuda)[None]
    graphed = RaftEngine(checkpoint=None, iters=4, seed=3, device=cuda, use_cuda_graph=True)
    eager = RaftEngine(checkpoint=None, iters=4, seed=3, device=cuda, use_cuda_graph=False)
    first = graphed.estimate_flow(a, b)
    keep = first.clone()
    second = graphed.estimate_flow(b, a)          # same graph key, different inputs
    assert first.data_ptr() != second.data_ptr()
    assert torch.equal(first, keep), 'the first result was overwritten by the next replay'
    assert first.shape == (1, 132, 160, 2)
    assert torch.allclose(first, eager.estimate_flow(a, b), atol=1e-4)
    assert not torch.allclose(first, second, atol=1e-2)


def test_keyed_flow_equals_pairwise_flow(cuda):
    """estimate_flow_keyed(key, frames) == estimate_flow(frames, key repeated): fnet(key) and the pooled correlation operands
    computed once per key (InstanceNorm is per image, so the key's features do not depend on the batch it is encoded in)."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    canvas = gi.texture(128 + 32, 160 + 32, 321)
    frames = np.stack([canvas[4 * i:4 * i + 128, 3 * i:3 * i + 160] for i in range(4)])
    key2 = gi.texture(128, 160, 322)
    fr = torch.from_numpy(frames).to(cuda)
    for graph in (False, True):
        eng = RaftEngine(checkpoint=None, iters=4, seed=0, device=cuda, use_cuda_graph=graph)
        k = eng.encode_key(fr[0])
        flow_k = eng.estimate_flow_keyed(k, fr[1:])
        flow_p = eng.estimate_flow(fr[1:], fr[:1].expand(3, -1, -1, -1).contiguous())
        assert flow_k.shape == flow_p.shape == (3, 128, 160, 2)
        d = (flow_k - flow_p).norm(dim=-1)
        print(f'keyed vs pairwise (graph={graph}): EPE mean {float(d.mean()):.2e} max {float(d.max()):.2e}')
        assert float(d.max()) <= 1e-3
        # a second key through the same graph (static operand buffers are refreshed), then the first one again
        k2 = eng.encode_key(torch.from_numpy(key2).to(cuda))
        f2 = eng.estimate_flow_keyed(k2, fr[1:])
        p2 = eng.estimate_flow(fr[1:], torch.from_numpy(key2).to(cuda)[None].expand(3, -1, -1, -1).contiguous())
        assert float((f2 - p2).norm(dim=-1).max()) <= 1e-3
        assert torch.allclose(eng.estimate_flow_keyed(k, fr[1:]), flow_k, atol=1e-5)


def test_sequence_flow_equals_pairwise_flow(cuda):
    """estimate_flow_sequence(frames) == estimate_flow(frames[:-1], frames[1:]): the feature encoder runs once per frame
    (InstanceNorm is per image, so a frame's features do not depend on the batch it is encoded in)."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    canvas = gi.texture(128 + 32, 160 + 32, 777)
    frames = np.stack([canvas[5 * i:5 * i + 128, 4 * i:4 * i + 160] for i in range(5)])
    fr = torch.from_numpy(frames).to(cuda)
    for graph in (False, True):
        eng = RaftEngine(checkpoint=None, iters=4, seed=0, device=cuda, use_cuda_graph=graph)
        flow_s = eng.estimate_flow_sequence(fr)
        flow_p = eng.estimate_flow(fr[:-1].contiguous(), fr[1:].contiguous())
        assert flow_s.shape == flow_p.shape == (4, 128, 160, 2)
        d = (flow_s - flow_p).norm(dim=-1)
        print(f'sequence vs pairwise (graph={graph}): EPE mean {float(d.mean()):.2e} max {float(d.max()):.2e}')
        assert float(d.max()) <= 1e-3
        # same graph, other frames; and the first result was a copy
        flow_r = eng.estimate_flow_sequence(fr.flip(0).contiguous())
        assert flow_r.data_ptr() != flow_s.data_ptr()
        assert float((flow_r - eng.estimate_flow(fr.flip(0)[:-1].contiguous(), fr.flip(0)[1:].contiguous())).norm(dim=-1).max()) <= 1e-3
    with pytest.raises(RuntimeError):
        eng.estimate_flow_sequence(fr[:1])


def test_deferred_coords_update_is_bit_identical(cuda):
    """The flow head's coords update applied inside the next lookup / convf1 (sdof_corr_lookup_gather_h,
    sdof_conv7x7_c2_relu_coords_h) instead of by its own kernel: same summation order, so the flow must not change by a bit."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    canvas = gi.texture(136 + 24, 152 + 24, 99)
    a = torch.from_numpy(np.stack([canvas[:136, :152], canvas[7:143, 3:155]])).to(cuda)
    b = torch.from_numpy(np.stack([canvas[4:140, 6:158], canvas[2:138, 9:161]])).to(cuda)
    torch.backends.cudnn.allow_tf32 = True        # the fp16 update loop (the only one with the deferred update) needs the TF32 default
    flows = []
    for defer in (True, False):
        eng = RaftEngine(checkpoint=None, iters=5, seed=1, device=cuda, use_cuda_graph=False, fast_options=dict(defer_coords=defer, convf1_gemm=False))
        assert eng.fast.loop_fp16 and eng.fast.defer_coords == defer
        flows.append(eng.estimate_flow(a, b))
    assert flows[0].shape == (2, 136, 152, 2)
    assert torch.equal(flows[0], flows[1])
    assert float(flows[0].abs().max()) > 0.1


def test_deferred_coords_kernels_vs_separate_update(cuda):
    """The two consumers of the deferred update against the separate kernel on the same taps: coordinates, flow, lookup rows and
    convf1 output identical."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(5)
    B, h, w = 2, 13, 21
    f1 = torch.randn((B, h, w, 64), generator=g, device=cuda)
    f2 = torch.randn((B, h, w, 64), generator=g, device=cuda)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, 'fp16', 'fp16')
    ys, xs = torch.meshgrid(torch.arange(h, device=cuda), torch.arange(w, device=cuda), indexing='ij')
    grid = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1).contiguous()
    cin = (grid + 2 * torch.randn((B, h, w, 2), generator=g, device=cuda)).contiguous()
    taps = torch.randn((B * h * w * 18,), generator=g, device=cuda) * 0.3
    bias = (0.25, -0.5)
    wT = torch.randn((7, 7, 2, 128), generator=g, device=cuda) * 0.1
    b7 = torch.randn((128,), generator=g, device=cuda) * 0.1
    # reference: separate update kernel, then the plain consumers
    c_ref, fl_ref = cin.clone(), torch.empty_like(cin)
    ops.flowhead2_gather_update(taps, bias, c_ref, fl_ref)
    look_ref = ops.corr_lookup_nhwc_h(pyr, c_ref, torch.empty((B, h, w, 328), dtype=torch.float16, device=cuda))
    conv_ref = ops.conv7x7_c2_relu_h(fl_ref, wT, b7, torch.empty((B, h, w, 128), dtype=torch.float16, device=cuda))
    # deferred
    cout, fl = torch.empty_like(cin), torch.empty_like(cin)
    look = ops.corr_lookup_gather_nhwc_h(pyr, cin, taps, bias, cout, fl, torch.empty((B, h, w, 328), dtype=torch.float16, device=cuda))
    conv = ops.conv7x7_c2_relu_coords_h(cin, taps, bias, wT, b7, torch.empty((B, h, w, 128), dtype=torch.float16, device=cuda))
    assert torch.equal(cout, c_ref) and torch.equal(fl, fl_ref)
    assert torch.equal(look, look_ref) and torch.equal(conv, conv_ref)
    # no taps: coordinates pass through
    look0 = ops.corr_lookup_gather_nhwc_h(pyr, cin, None, bias, cout, fl, torch.empty((B, h, w, 328), dtype=torch.float16, device=cuda))
    assert torch.equal(cout, cin) and torch.equal(fl, cin - grid)
    assert torch.equal(look0, ops.corr_lookup_nhwc_h(pyr, cin, torch.empty((B, h, w, 328), dtype=torch.float16, device=cuda)))
    with pytest.raises(RuntimeError):
        ops.corr_lookup_gather_nhwc_h(pyr, cin, taps, bias, cin, fl, look)


def test_convf1_as_im2col_gemm_matches_the_fma_kernel(cuda):
    """convf1 as im2col rows (sdof_flow_im2col7_h, fp16 hi/lo split of the flow) + a 1x1 fp16 convolution vs the fp32 FMA kernel:
    only the filter is rounded to fp16 (2^-11 relative per weight), large flows keep their low bits through the split."""
    import torch.nn.functional as F
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(11)
    B, h, w = 2, 13, 21
    ys, xs = torch.meshgrid(torch.arange(h, device=cuda), torch.arange(w, device=cuda), indexing='ij')
    grid = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1).contiguous()
    cin = (grid + 60 * torch.randn((B, h, w, 2), generator=g, device=cuda)).contiguous()       # flows of tens of pixels
    taps = torch.randn((B * h * w * 18,), generator=g, device=cuda) * 0.3
    bias = (0.25, -0.5)
    wt = torch.randn((128, 2, 7, 7), generator=g, device=cuda) * 0.1
    b7 = torch.randn((128,), generator=g, device=cuda) * 0.1
    ref = ops.conv7x7_c2_relu_coords_h(cin, taps, bias, wt.permute(2, 3, 1, 0).contiguous(), b7,
                                       torch.empty((B, h, w, 128), dtype=torch.float16, device=cuda)).float()
    rows = ops.flow_im2col7_h(cin, taps, bias, torch.empty((B, h, w, 200), dtype=torch.float16, device=cuda))
    # rows against the flow itself: hi + lo reproduces the fp32 flow, out-of-image taps are zero
    c_ref, fl_ref = cin.clone(), torch.empty_like(cin)
    ops.flowhead2_gather_update(taps, bias, c_ref, fl_ref)
    pad = F.pad(fl_ref.permute(0, 3, 1, 2), (3, 3, 3, 3))
    cols = F.unfold(pad, 7).view(B, 2, 49, h, w).permute(0, 3, 4, 2, 1).reshape(B, h, w, 98)    # [tap][ci]
    hi, lo = rows[..., :98].float(), rows[..., 98:196].float()
    assert float((hi + lo - cols).abs().max()) <= 2e-5 * float(cols.abs().max())
    assert float(rows[..., 196:].abs().max()) == 0.0
    y = torch.cudnn_convolution_relu(rows.permute(0, 3, 1, 2), ops.im2col7_weight(wt, 200), b7.half(), (1, 1), (0, 0), (1, 1), 1)
    y = y.permute(0, 2, 3, 1).float()
    err = float((y - ref).abs().max())
    print(f'convf1 im2col GEMM vs FMA kernel: max abs err {err:.3e} (max |out| {float(ref.abs().max()):.1f})')
    assert err <= 4e-3 * float(ref.abs().max())


def test_convf1_gemm_path_flow_stays_close(cuda):
    from sd_animation_optical_flow_b200.engine import RaftEngine
    canvas = gi.texture(136 + 24, 152 + 24, 98)
    a = torch.from_numpy(np.stack([canvas[:136, :152]])).to(cuda)
    b = torch.from_numpy(np.stack([canvas[4:140, 6:158]])).to(cuda)
    torch.backends.cudnn.allow_tf32 = True        # the fp16 update loop needs the TF32 default
    flows = [RaftEngine(checkpoint=None, iters=6, seed=2, device=cuda, use_cuda_graph=False,
                        fast_options=dict(convf1_gemm=gemm)).estimate_flow(a, b) for gemm in (True, False)]
    d = (flows[0] - flows[1]).norm(dim=-1)
    mag = float(flows[1].norm(dim=-1).mean())
    print(f'convf1 GEMM vs FMA path: EPE mean {float(d.mean()):.2e} max {float(d.max()):.2e}, mean |flow| {mag:.1f} px')
    # > 0: the two paths really differ (fp16-rounded filter); the random-init update block amplifies any rounding with the
    # flow magnitude, so the bar is relative like the runaway-weights bar of test_gpu_parity_full.py
    assert 0.0 < float(d.mean()) <= max(2e-3, 1e-3 * mag)


def test_engine_on_a_device_that_is_not_current(cuda):
    """engine, warp and masks on cuda:1 while cuda:0 is current (PDCNetAux(device=cuda:N))."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from sd_animation_optical_flow_b200 import ops
    from sd_animation_optical_flow_b200.engine import RaftEngine
    d1 = torch.device('cuda', 1)
    torch.cuda.set_device(0)
    f1, f2 = gi.raft_inputs('basic')
    e0 = RaftEngine(checkpoint=None, iters=3, seed=0, device=cuda)
    e1 = RaftEngine(checkpoint=None, iters=3, seed=0, device=d1)
    r0 = e0.estimate_flow(torch.from_numpy(f1).to(cuda)[None], torch.from_numpy(f2).to(cuda)[None])
    r1 = e1.estimate_flow(torch.from_numpy(f1).to(d1)[None], torch.from_numpy(f2).to(d1)[None])
    assert torch.cuda.current_device() == 0 and r1.device == d1
    assert torch.allclose(r0, r1.to(cuda), atol=1e-4)
    img = torch.from_numpy(f1).to(d1)
    w1 = ops.warp(img, r1[0])
    w0 = ops.warp(img.to(cuda), r1[0].to(cuda))
    assert torch.cuda.current_device() == 0 and torch.equal(w0, w1.to(cuda))


@pytest.mark.parametrize('shape', [(1, 96, 64), (2, 22, 40), (1, 11, 20), (1, 45, 80)])
def test_tcgen05_gru_kernels_vs_torch(cuda, shape):
    """csrc/conv_tc.cu (one SepConvGRU pass = gru_zr_tc + gru_q_tc, tap-shifted implicit GEMM on tcgen05 with the gate
    arithmetic in the epilogue) against F.conv2d in fp32 on the SAME fp16-rounded operands: what differs is the summation
    order and the fp16 rounding of r*h / h, so 2e-3 abs on O(1) gates (operand precision = TF32's 11 bits)."""
    import torch.nn.functional as F
    from sd_animation_optical_flow_b200 import ops
    B, h, w = shape
    g = torch.Generator(device=cuda).manual_seed(h)
    rnd = lambda *s: torch.randn(s, generator=g, device=cuda)
    for horizontal in (True, False):
        ks = (1, 5) if horizontal else (5, 1)
        pad = (0, 2) if horizontal else (2, 0)
        w_zr = rnd(384, 256, *ks) * 0.03
        w_q = rnd(128, 128, *ks) * 0.05
        H = torch.tanh(rnd(B, h, w, 128))
        hx16 = torch.cat([H, torch.relu(rnd(B, h, w, 126)), 3 * rnd(B, h, w, 2)], -1).half().contiguous()
        zrmap, qmap = rnd(B, h, w, 256) * 0.5, rnd(B, h, w, 128) * 0.5
        # reference in fp32 on the fp16-rounded operands
        x = hx16.float().permute(0, 3, 1, 2)
        zrq = F.conv2d(x, w_zr.half().float(), None, padding=pad).permute(0, 2, 3, 1)
        z_ref = torch.sigmoid(zrq[..., :128] + zrmap[..., :128])
        rh_ref = (torch.sigmoid(zrq[..., 128:256] + zrmap[..., 128:]) * H).half()
        qx_ref = zrq[..., 256:]
        q_ref = F.conv2d(rh_ref.float().permute(0, 3, 1, 2), w_q.half().float(), None, padding=pad).permute(0, 2, 3, 1)
        h_ref = (1 - z_ref) * H + z_ref * torch.tanh(q_ref + qx_ref + qmap)
        Z, QX = torch.empty_like(H), torch.empty_like(H)
        RH16 = torch.empty((B, h, w, 128), dtype=torch.float16, device=cuda)
        Hc = H.clone()
        ops.gru_zr_tc(hx16, ops.gru_weights16(w_zr), zrmap, Hc, horizontal, Z, RH16, QX)
        assert float((Z - z_ref).abs().max()) <= 1e-3
        assert float((QX - qx_ref).abs().max()) <= 2e-3
        assert float((RH16.float() - rh_ref.float()).abs().max()) <= 2e-3
        keep_tail = hx16[..., 128:].clone()
        ops.gru_q_tc(RH16, ops.gru_weights16(w_q), qmap, QX, Z, horizontal, Hc, hx16)
        assert float((Hc - h_ref).abs().max()) <= 3e-3
        assert torch.equal(hx16[..., :128], Hc.half()) and torch.equal(hx16[..., 128:], keep_tail)
    # motion-encoder tail -> fp16 GRU input
    mc, mf, bias, flow = rnd(B, h, w, 128), rnd(B, h, w, 128), rnd(128), 5 * rnd(B, h, w, 2)
    hx = torch.zeros((B, h, w, 256), dtype=torch.float16, device=cuda)
    ops.motion_tail16(mc, mf, bias, flow, hx)
    want = torch.cat([torch.relu(mc + mf + bias)[..., :126], flow], -1).half()
    assert torch.equal(hx[..., 128:], want) and float(hx[..., :128].abs().max()) == 0


def test_tcgen05_gru_path_equals_cudnn_gru_path(cuda):
    """FastRaft with the tensor-core GRU (opt-in, fast_options=dict(tc_gru=True)) against the default forward with cuDNN
    convolutions (TF32 off in this module) + glue kernels."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    f1, f2 = gi.shifted_pair(128, 160, 55)
    a = torch.from_numpy(f1).to(cuda)[None]
    b = torch.from_numpy(f2).to(cuda)[None]
    tc = RaftEngine(checkpoint=None, iters=8, seed=0, flow_head_scale=0.02, device=cuda, use_cuda_graph=False, fast_options=dict(tc_gru=True))
    ref = RaftEngine(checkpoint=None, iters=8, seed=0, flow_head_scale=0.02, device=cuda, use_cuda_graph=False)
    assert tc.fast.tc_gru and not ref.fast.tc_gru
    d = (tc.estimate_flow(a, b) - ref.estimate_flow(a, b)).norm(dim=-1)
    print(f'tcgen05 GRU vs cuDNN fp32 GRU: EPE mean {float(d.mean()):.2e} max {float(d.max()):.2e}')
    assert float(d.mean()) <= 2e-3 and float(d.max()) <= 2e-2


def test_fp16_feature_encoder_matches_fp32_encoder(cuda):
    """FastEncoder(fnet, fp16) -- fp16 activations / filters, fp32 accumulation, fp32/fp64 normalisation statistics -- against the
    fp32 encoder on the same weights: 5e-3 of the feature range (15 conv + norm layers of 2^-11 operand rounding each)."""
    from sd_animation_optical_flow_b200.engine import RaftEngine
    from sd_animation_optical_flow_b200 import ops
    eng = RaftEngine(checkpoint=None, iters=2, seed=0, device=cuda, use_cuda_graph=False)
    assert eng.fast._fnet16 is not None
    f1, _ = gi.shifted_pair(128, 160, 7)
    im = ops.normalize_pad_u8(torch.from_numpy(f1).to(cuda)[None], (0, 0, 0, 0), channels=4)
    a = eng.fast._fnet32(im)
    b = eng.fast._fnet16(im)
    assert a.dtype == b.dtype == torch.float32 and a.shape == b.shape
    rel = float((a - b).abs().max() / a.abs().max())
    print(f'fp16 vs fp32 feature encoder: max rel diff {rel:.2e}')
    assert rel <= 5e-3


def test_fp16_glue_kernels_vs_fp32_glue(cuda):
    """csrc/raft_glue16.cu against the fp32 glue kernels on the same (fp16-representable) inputs: results agree to fp16 output
    rounding (2^-11 relative)."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(11)
    rnd = lambda *s: torch.randn(s, generator=g, device=cuda)
    B, h, w = 2, 13, 21
    npix = B * h * w
    H = torch.tanh(rnd(B, h, w, 128))
    zr16 = rnd(B, h, w, 384).half()
    q16 = rnd(B, h, w, 128).half()
    zrmap, qmap = rnd(B, h, w, 256), rnd(B, h, w, 128)
    # gru_rh
    rh32 = torch.empty((B, h, w, 128), device=cuda)
    ops.gru_rh(zr16.float(), H, rh32, bias_zr=zrmap)
    rh16 = torch.empty((B, h, w, 128), device=cuda, dtype=torch.float16)
    ops.gru_rh_h(zr16, zrmap, H, rh16)
    assert float((rh16.float() - rh32).abs().max()) <= 1e-3
    # gru_update
    H32, HX32 = H.clone(), torch.zeros((B, h, w, 256), device=cuda)
    ops.gru_update(zr16.float(), q16.float(), H32, HX32, bias_zr=zrmap, bias_q=qmap)
    Hh, HX16, H16 = H.clone(), torch.zeros((B, h, w, 256), device=cuda, dtype=torch.float16), torch.zeros((B, h, w, 128), device=cuda, dtype=torch.float16)
    ops.gru_update_h(zr16, zrmap, q16, qmap, Hh, HX16, H16)
    assert float((Hh - H32).abs().max()) <= 1e-5
    assert torch.equal(HX16[..., :128], Hh.half()) and torch.equal(H16, Hh.half()) and float(HX16[..., 128:].abs().max()) == 0
    # motion tail
    mc, mf, bias, flow = rnd(B, h, w, 128).half(), rnd(B, h, w, 128).half(), rnd(128), 5 * rnd(B, h, w, 2)
    hx = torch.zeros((B, h, w, 256), dtype=torch.float16, device=cuda)
    ops.motion_tail16_h(mc, mf, bias, flow, hx)
    want = torch.cat([torch.relu(mc.float() + mf.float() + bias)[..., :126], flow], -1).half()
    assert torch.equal(hx[..., 128:], want)
    # conv7x7 -> fp16, flow-head taps from fp16
    wt = (rnd(7, 7, 2, 128) * 0.1).contiguous()
    b7 = rnd(128)
    o32 = ops.conv7x7_c2_relu(flow, wt, b7)
    o16 = torch.empty((B, h, w, 128), device=cuda, dtype=torch.float16)
    ops.conv7x7_c2_relu_h(flow, wt, b7, o16)
    assert torch.equal(o16, o32.half())
    x16 = torch.relu(rnd(B, h, w, 256)).half()
    w2 = (rnd(3, 3, 2, 256) * 0.05).contiguous()
    ys, xs = torch.meshgrid(torch.arange(h, device=cuda), torch.arange(w, device=cuda), indexing='ij')
    grid = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1)
    c_a, c_b = (grid + 0.5).contiguous(), (grid + 0.5).contiguous()
    f_a, f_b = torch.empty((B, h, w, 2), device=cuda), torch.empty((B, h, w, 2), device=cuda)
    ops.flowhead2_update(x16.float(), w2, (0.3, -0.7), c_a, f_a, None, 0, None, 0)
    ops.flowhead2_update_h(x16, w2, (0.3, -0.7), c_b, f_b, torch.empty((npix * 18,), device=cuda))
    assert float((c_a - c_b).abs().max()) <= 1e-5 and float((f_a - f_b).abs().max()) <= 1e-5
    # lookup -> fp16, padded channels
    from sd_animation_optical_flow_b200.raft import coords_grid
    f1, f2 = rnd(1, 16, 24, 64), rnd(1, 16, 24, 64)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, 'fp16', 'fp16')
    cn = (coords_grid(1, 16, 24, cuda) + 2 * rnd(1, 2, 16, 24)).permute(0, 2, 3, 1).contiguous()
    l32 = torch.empty((1, 16, 24, 324), device=cuda)
    ops.corr_lookup_nhwc(pyr, cn, 4, l32)
    l16 = torch.full((1, 16, 24, 328), 7.0, device=cuda, dtype=torch.float16)
    ops.corr_lookup_nhwc_h(pyr, cn, l16)
    assert torch.equal(l16[..., :324], l32.half()) and float(l16[..., 324:].abs().max()) == 0
