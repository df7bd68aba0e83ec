"""N2: ParticleSystem.Render (RasterizeParticleSystem.fx).  CPU part: known answers for the oracle restatement and the host
packing of Uniforms.RasterizeParticleSystem.  GPU part (marked): the tiled CUDA rasteriser against the oracle through
ilb_particles_render, plus a size-independent checksum property at 4K with a million particles."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi

f32 = np.float32


def _system(ctx, chunk=32, size=(1.0, 1.0), max_chunks=4, **appearance):
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=3))
    cfg = ib.ParticleSystemConfiguration()
    cfg.Size = size
    cfg.Appearance = ib.ParticleAppearance(**appearance)
    return ib.ParticleSystem(engine, cfg, maxChunks=max_chunks)


def _sprite_sheet(w, h, seed=1):
    rs = np.random.RandomState(seed)
    t = rs.randint(0, 256, size=(h, w, 4), dtype=np.uint8)
    t[..., 3] = np.where(rs.rand(h, w) < 0.15, 0, t[..., 3] | 128)
    t[..., :3] = (t[..., :3].astype(np.float32) * (t[..., 3:4].astype(np.float32) / 255)).astype(np.uint8)    # premultiplied
    return t


def _random_state(n, w, h, seed, size_range=(1.0, 9.0), dead=0.1):
    rs = np.random.RandomState(seed)
    P = np.zeros((n, 4), np.float32)
    P[:, 0], P[:, 1], P[:, 2] = rs.rand(n) * (w + 40) - 20, rs.rand(n) * (h + 40) - 20, rs.rand(n) * 10
    P[:, 3] = np.where(rs.rand(n) < dead, 0.0, rs.rand(n) * 4 + 0.1)
    RD = np.zeros((n, 4), np.float32)
    RD[:, 0] = rs.rand(n) * (size_range[1] - size_range[0]) + size_range[0]
    RD[:, 1] = rs.rand(n) * 20 - 3            # rotation, several turns, negative too (fmod keeps the sign)
    RD[:, 2] = rs.rand(n) * 50
    RD[:, 3] = rs.randint(0, 4, n)            # category -> sprite row
    RC = rs.rand(n, 4).astype(np.float32)
    RC[:, :3] *= RC[:, 3:4]
    RC[rs.rand(n) < 0.05] = 0                 # fully transparent: discarded
    return P, RD, RC


# ------------------------------------------------------------------------------------------------------------- CPU / oracle
def test_render_params_packing():
    tex = _sprite_sheet(64, 32)
    s = _system(None, size=(2.0, 3.0), Texture=tex, OffsetPx=(16, 0), SizePx=(16, 8), AnimationRate=(4.0, -2.0), Rounded=True,
                Bilinear=False, ColumnFromVelocity=True)
    s.Configuration.GlobalColor = (1.0, 0.5, 0.25, 0.5)
    s.Configuration.SizeFromZ = 0.1
    r = s.render_params(640, 480, "Additive", ib.ParticleRenderParameters(Origin=(5, 6), Scale=(2, 2)), (10, 20), (1.5, 1.5))
    assert r.BitmapTextureRegion.tuple() == (0.25, 0.0, 0.5, 0.25)                    # Uniforms.cs:262-270
    assert r.SizeFactorAndPosition.tuple() == (8.0, 4.0, 5.0, 6.0)                    # RelativeSize: SizePx * 0.5 (:274-275)
    assert r.GlobalColor.tuple() == (0.5, 0.25, 0.125, 0.5)                           # premultiplied (:281-285)
    assert (r.texture_filter, r.blend, r.clear) == (_abi.TEXTURE_POINT, _abi.BLEND_ADDITIVE, 0)
    assert r.RenderingOptions.tuple() == (1, 0, 1, 0) and r.ZConfiguration.x == pytest.approx(0.1)
    assert r.TexelAndSize.tuple() == (1 / 32, 1 / 32, 2.0, 3.0) and r.RoundingPowerFromLife.ABCD.x == pytest.approx(0.8)
    assert (r.AnimationRateAndRotationAndZToY.x, r.AnimationRateAndRotationAndZToY.y) == (0.25, -0.5)
    assert tuple(r.ViewportPosition) == (10.0, 20.0) and tuple(r.ViewportScale) == (1.5, 1.5) and r.StippleFactor == 1.0
    s2 = _system(None)
    r2 = s2.render_params(8, 8, clearColor=(0, 0, 0, 1))
    assert r2.texture_filter == _abi.TEXTURE_NONE and r2.BitmapTextureRegion.tuple() == (0, 0, 1, 1) and r2.clear == 1


def test_oracle_quads_blend_in_draw_order(oracle):
    s = _system(None, chunk=8, size=(4.0, 2.0))
    r = s.render_params(32, 24, clearColor=(0, 0, 0, 0))
    P, RD, RC = (np.zeros((3, 4), np.float32) for _ in range(3))
    P[0], RD[0], RC[0] = [16, 12, 0, 1], [1.0, 0.0, 0, 0], [0.5, 0.25, 0.125, 0.5]
    P[1], RD[1], RC[1] = [18, 12, 0, 1], [1.0, np.pi / 2, 0, 0], [0.2, 0.2, 0.2, 1.0]      # quarter turn: 4 x 8 instead of 8 x 4
    P[2], RD[2], RC[2] = [16, 12, 0, 0], [5.0, 0.0, 0, 0], [1, 1, 1, 1]                    # dead: draws nothing
    out = oracle.particles_render(P, RD, RC, r)
    cov = out[..., 3] > 0
    assert cov.sum() == 48 and cov[10:14, 12:20].all() and cov[8:16, 16:20].all()
    assert np.allclose(out[12, 13], [0.5, 0.25, 0.125, 0.5]) and np.allclose(out[12, 17], [0.2, 0.2, 0.2, 1.0])   # opaque quad on top
    out = oracle.particles_render(P[[1, 0, 2]], RD[[1, 0, 2]], RC[[1, 0, 2]], r)           # reversed order: premultiplied "over"
    assert np.allclose(out[12, 17], np.array([0.5, 0.25, 0.125, 0.5]) + np.array([0.2, 0.2, 0.2, 1.0]) * 0.5)
    r.blend = _abi.BLEND_ADDITIVE                                                           # src * src.a + dst
    out = oracle.particles_render(P, RD, RC, r)
    assert np.allclose(out[12, 17], np.array([0.5, 0.25, 0.125, 0.5]) * 0.5 + np.array([0.2, 0.2, 0.2, 1.0]))
    r.blend, r.clear = _abi.BLEND_ALPHA, 0                                                  # over existing contents
    base = np.full((24, 32, 4), 0.25, np.float32)
    out = oracle.particles_render(P, RD, RC, r, target=base)
    assert np.allclose(out[0, 0], 0.25) and np.allclose(out[12, 13], np.array([0.5, 0.25, 0.125, 0.5]) + 0.25 * 0.5)


def test_oracle_rounded_scaled_and_textured(oracle):
    s = _system(None, chunk=8, size=(8.0, 8.0), Rounded=True)
    r = s.render_params(64, 64, "AlphaBlend", ib.ParticleRenderParameters(Origin=(4, 0), Scale=(2, 1)), (10, 5), (1.0, 2.0), clearColor=(0, 0, 0, 0))
    P, RD, RC = (np.zeros((1, 4), np.float32) for _ in range(3))
    P[0], RD[0], RC[0] = [20, 15, 5, 1], [1.0, 0.0, 0, 0], [1, 1, 1, 1]
    s.Configuration.ZToY = 0.0
    out = oracle.particles_render(P, RD, RC, r)
    # centre = ((20*2 + 4) - 10) * 1, ((15*1 + 0) - 5) * 2 = (34, 20); half extents 8*2*1 = 16 and 8*1*2 = 16
    ys, xs = np.nonzero(out[..., 3] > 0)
    assert abs(xs.mean() + 0.5 - 34) < 0.01 and abs(ys.mean() + 0.5 - 20) < 0.01
    assert out[20, 34, 3] == 1.0 and out[5, 19, 3] == 0.0                  # centre opaque, corner rounded away
    assert xs.min() >= 18 and xs.max() <= 49 and 0.7 < (out[..., 3] > 0).sum() / (32 * 32) < 0.95
    # a 2 x 2 sprite sheet (POINT): rows from the category (renderData.w), columns from life * AnimationRate
    tex = np.zeros((4, 4, 4), np.uint8)
    for fy in range(2):
        for fx in range(2):
            tex[2 * fy:2 * fy + 2, 2 * fx:2 * fx + 2] = [60 * (fx + 1), 60 * (fy + 1), 0, 255]
    st = _system(None, chunk=8, size=(1.0, 1.0), Texture=tex, SizePx=(2, 2), AnimationRate=(1.0, 0.0), Bilinear=False, RelativeSize=False)
    rt = st.render_params(16, 16, clearColor=(0, 0, 0, 0))
    P2, RD2, RC2 = (np.zeros((4, 4), np.float32) for _ in range(3))
    for k, (life, cat) in enumerate([(0.5, 0), (1.5, 0), (0.5, 1), (1.5, 1)]):
        P2[k], RD2[k], RC2[k] = [2 + 4 * k, 2, 0, life], [1.0, 0.0, 0, cat], [1, 1, 1, 1]
    out = oracle.particles_render(P2, RD2, RC2, rt, texture=tex)
    got = [tuple(np.round(out[2, 2 + 4 * k, :2] * 255).astype(int)) for k in range(4)]
    assert got == [(60, 60), (120, 60), (60, 120), (120, 120)]


# ------------------------------------------------------------------------------------------------------------------- GPU
def _upload(system, P, RD, RC):
    """Puts (P, RD, RC) into the system's PositionAndLife / RenderData / RenderColor: Spawn fills P (velocity 0, attributes =
    RC), then RenderColor / RenderData -- outputs of the update pass -- are written with ilb_particles_upload_buffer."""
    n = P.shape[0]
    system.Spawn(P, np.zeros_like(P), RC)
    per = system.ChunkMaximumCount
    total = system.LiveChunkCount * per
    Pf, RDf, RCf = (np.zeros((total, 4), np.float32) for _ in range(3))
    Pf[:n], RDf[:n], RCf[:n] = P, RD, RC
    for c in range(system.LiveChunkCount):
        system.WriteChunkBuffer(c, 3, RCf[c * per:(c + 1) * per])
        system.WriteChunkBuffer(c, 4, RDf[c * per:(c + 1) * per])
    return Pf, RDf, RCf


def _compare(gpu, ref, what, exact_coverage=True):
    assert gpu.shape == ref.shape and not np.isnan(gpu).any(), what
    if exact_coverage:
        assert np.array_equal(gpu[..., 3] != 0, ref[..., 3] != 0), f"{what}: coverage differs"
    # premultiplied colours in [0, 1]: 1e-4 relative with an absolute floor of 1e-6 (1/250 of a Color LSB).  Coverage, geometry and
    # blending are bit-identical by construction; the only inexact operation is powf in computeCircularAlpha (CUDA vs glibc).
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64)) / np.maximum(np.abs(ref.astype(np.float64)), 1e-2)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= 1e-4, f"{what}: max rel err {err.max():.3e} at {worst}: gpu {gpu[worst]} ref {ref[worst]}"


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,blend,rounded", [(96, 64, 700, "AlphaBlend", False), (101, 67, 1500, "Additive", True),
                                                  (33, 17, 4000, "AlphaBlend", True), (256, 160, 3000, "Opaque", False)])
def test_gpu_render_untextured(ctx, oracle, w, h, n, blend, rounded):
    s = _system(ctx, size=(1.5, 1.0), Rounded=rounded)
    s.Configuration.GlobalColor = (0.9, 0.8, 1.0, 0.75)
    s.Configuration.SizeFromZ = 0.05
    s.Configuration.ZToY = 0.5
    P, RD, RC = _random_state(n, w, h, seed=n)       # 33x17 with 4000 quads: > 256 quads per tile (several shared-memory batches)
    Pf, RDf, RCf = _upload(s, P, RD, RC)
    r = s.render_params(w, h, blend, ib.ParticleRenderParameters(Origin=(1.5, -2.0), Scale=(1.25, 0.75)), (3.0, 1.0), (1.1, 1.3), clearColor=(0.1, 0.0, 0.2, 0.0))
    ref = oracle.particles_render(Pf, RDf, RCf, r)
    gpu = s.Render(w, h, None, blend, ib.ParticleRenderParameters(Origin=(1.5, -2.0), Scale=(1.25, 0.75)), (3.0, 1.0), (1.1, 1.3), (0.1, 0.0, 0.2, 0.0))
    _compare(gpu, ref, f"untextured {w}x{h} {blend}", exact_coverage=(blend != "Additive" and not rounded))


@pytest.mark.gpu
@pytest.mark.parametrize("bilinear,relative,column,row", [(True, True, False, False), (False, True, True, False), (True, False, False, True)])
def test_gpu_render_textured_sprite_sheet(ctx, oracle, bilinear, relative, column, row):
    tex = _sprite_sheet(64, 48, seed=4)
    s = _system(ctx, size=(0.5, 0.5) if relative else (6.0, 5.0), Texture=tex, OffsetPx=(0, 0), SizePx=(16, 12), AnimationRate=(0.7, -1.3),
                Bilinear=bilinear, RelativeSize=relative, ColumnFromVelocity=column, RowFromVelocity=row, Rounded=row)
    P, RD, RC = _random_state(1200, 160, 120, seed=9, size_range=(0.5, 2.0))
    Pf, RDf, RCf = _upload(s, P, RD, RC)
    base = np.random.RandomState(2).rand(120, 160, 4).astype(np.float32)
    r = s.render_params(160, 120, "AlphaBlend")
    ref = oracle.particles_render(Pf, RDf, RCf, r, texture=tex, target=base)
    gpu = s.Render(160, 120, base, "AlphaBlend")
    _compare(gpu, ref, f"textured bilinear={bilinear}", exact_coverage=False)
    assert np.abs(gpu - base).max() > 0.1


@pytest.mark.gpu
def test_gpu_render_formats_big_quads_and_empty(ctx, oracle):
    s = _system(ctx, size=(1.0, 1.0))
    empty = s.Render(40, 24, None, clearColor=(0.25, 0.5, 0.75, 1.0))                 # no chunks at all: the clear colour
    assert np.all(empty == np.array([0.25, 0.5, 0.75, 1.0], np.float32))
    P, RD, RC = _random_state(300, 200, 150, seed=21, size_range=(20.0, 90.0))       # quads that span dozens of tiles
    Pf, RDf, RCf = _upload(s, P, RD, RC)
    r = s.render_params(200, 150, clearColor=(0, 0, 0, 0))
    ref = oracle.particles_render(Pf, RDf, RCf, r)
    _compare(s.Render(200, 150, None), ref, "big quads")
    half = s.Render(200, 150, np.zeros((150, 200, 4), np.float16))
    assert half.dtype == np.float16 and np.abs(half.astype(np.float32) - ref).max() <= 2e-3 * max(1.0, float(np.abs(ref).max()))
    rgba = s.Render(200, 150, np.zeros((150, 200, 4), np.uint8))
    want = np.floor(np.clip(ref, 0, 1) * 255 + 0.5).astype(np.int32)
    assert rgba.dtype == np.uint8 and np.abs(rgba.astype(np.int32) - want).max() <= 1
    with pytest.raises(ib.IlluminantError) as e:
        s.Render(200, 150, None, renderParams=ib.ParticleRenderParameters(StippleFactor=0.5))
    assert e.value.code == _abi.ERR_UNSUPPORTED
    s.Configuration.Appearance.DitheredOpacity = True
    with pytest.raises(ib.IlluminantError):
        s.Render(200, 150, None)


@pytest.mark.gpu
def test_gpu_render_checksum_property_at_4k(ctx):
    """Size-independent property at BASELINE.json's frame size: a million live particles drawn additively as axis-aligned
    one-pixel white quads at pixel centres each add exactly 1 to exactly one pixel, whatever tile they fall into: the target's sum
    is the number of on-screen particles and its per-pixel values are the particle counts per pixel."""
    w, h, n = 3840, 2160, 1 << 20
    s = _system(ctx, chunk=512, size=(0.5, 0.5), max_chunks=4)
    rs = np.random.RandomState(5)
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = rs.randint(-50, w + 50, n) + 0.5
    P[:, 1] = rs.randint(-50, h + 50, n) + 0.5
    P[:, 3] = 1.0
    RD = np.zeros((n, 4), np.float32)
    RD[:, 0] = 1.0
    RC = np.ones((n, 4), np.float32)
    _upload(s, P, RD, RC)
    out = s.Render(w, h, None, "Additive")
    xi, yi = np.floor(P[:, 0]).astype(int), np.floor(P[:, 1]).astype(int)
    on = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
    counts = np.zeros((h, w), np.float32)
    np.add.at(counts, (yi[on], xi[on]), 1.0)
    assert np.array_equal(out[..., 0], counts) and np.array_equal(out[..., 3], counts)
    assert float(out[..., 1].sum(dtype=np.float64)) == float(on.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("blend", ["AlphaBlend", "Additive"])
def test_gpu_layers_of_chunk_ranges_composite_to_the_whole_render(ctx, oracle, blend):
    """Multi-GPU ParticleSystem.Render on one device: the chunks of a system are split into three contiguous ranges ("ranks"),
    each range is rendered over a transparent float4 layer, and the layers are composited in range order -- band by band, as the
    ranks would -- onto the clear colour.  Premultiplied "over" is associative and additive blending is a sum, so the result is
    the single render of all chunks up to fp32 reassociation; the oracle (one pass over all particles) is the reference."""
    import torch
    from illuminant_b200.particles import composite_layers
    w, h, n, chunk = 160, 96, 2700, 16
    params = ib.ParticleRenderParameters(Origin=(0.5, 0.25), Scale=(1.1, 0.9))
    clear = (0.05, 0.1, 0.0, 0.2)
    P, RD, RC = _random_state(n, w, h, seed=17, size_range=(2.0, 12.0))
    whole = _system(ctx, chunk=chunk, max_chunks=12, Rounded=True)
    Pf, RDf, RCf = _upload(whole, P, RD, RC)
    nchunks, per = whole.LiveChunkCount, chunk * chunk
    assert nchunks == 11
    ref = oracle.particles_render(Pf, RDf, RCf, whole.render_params(w, h, blend, params, clearColor=clear))
    single = whole.Render(w, h, None, blend, params, clearColor=clear)
    _compare(single, ref, f"single render {blend}", exact_coverage=False)
    layers, systems = [], []
    for c0, c1 in ((0, 4), (4, 8), (8, 11)):                         # sharding.chunk_range(rank, 3, 11)
        part = _system(ctx, chunk=chunk, max_chunks=4, Rounded=True)
        sl = slice(c0 * per, c1 * per)
        _upload(part, Pf[sl], RDf[sl], RCf[sl])
        layer = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
        part.RenderLayerDevice(layer.data_ptr(), w, h, blend, params)
        layers.append(layer)
        systems.append(part)
    target = torch.full((h, w, 4), 7.0, dtype=torch.float32, device="cuda")
    for r0, r1 in ((0, 32), (32, 64), (64, 96)):                      # each "rank" composites its own row band of every layer
        composite_layers(ctx, [l.data_ptr() for l in layers], w, h, (r0, r1), blend, _abi.FORMAT_FLOAT4, clear, [target.data_ptr()])
    ctx.synchronize()
    got = target.cpu().numpy()
    _compare(got, ref, f"composited layers {blend}", exact_coverage=False)
    assert np.abs(got - single).max() <= 2e-6                         # the same image as the one-GPU render, up to reassociation
    # over an existing target (clear_color = NULL) and into a half4 target
    base = torch.rand((h, w, 4), dtype=torch.float32, device="cuda")
    expect = base.clone()
    for l in layers:
        expect = l + expect * (1.0 - l[..., 3:4]) if blend == "AlphaBlend" else l + expect
    composite_layers(ctx, [l.data_ptr() for l in layers], w, h, (0, h), blend, _abi.FORMAT_FLOAT4, None, [base.data_ptr()])
    ctx.synchronize()
    assert torch.allclose(base, expect, rtol=0, atol=2e-6)
    half = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
    composite_layers(ctx, [l.data_ptr() for l in layers], w, h, (0, h), blend, _abi.FORMAT_HALF4, clear, [half.data_ptr()])
    ctx.synchronize()
    assert torch.equal(half, target.half())
    with pytest.raises(ib.IlluminantError):
        composite_layers(ctx, [l.data_ptr() for l in layers], w, h, (0, h), "Opaque", _abi.FORMAT_FLOAT4, clear, [target.data_ptr()])
