"""GPU parity of the lighting hot path (L1-L11) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes
from helpers import LIGHTING_RTOL, lighting_rel_err, make_renderer, oracle_lightmap

pytestmark = pytest.mark.gpu


def _check(gpu, ref, what):
    err = lighting_rel_err(gpu, ref)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= LIGHTING_RTOL, f"{what}: max rel err {err.max():.3e} at {worst}: gpu {gpu[worst]} ref {ref[worst]}"


def test_c1_single_sphere_light(ctx, oracle):
    s = scenes.config_c1()
    s.configuration.Float4Lightmap = True
    r, tex = make_renderer(ctx, s)
    _check(r.RenderLighting(), oracle_lightmap(oracle, r, tex, s), "C1")


def test_df_generation_matches_oracle(ctx, oracle):
    s = scenes.lighting_scene(11, 320, 200, 0)
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    gpu = df.Save().astype(np.int32)
    ref = oracle.generate_distance_field(df, s.obstructions).astype(np.int32)
    assert np.abs(gpu - ref).max() <= 1          # UNORM16 LSB
    assert (gpu != ref).mean() < 1e-3


@pytest.mark.parametrize("seed,w,h,ns,nd,nl", [(21, 320, 200, 6, 0, 0), (22, 257, 131, 3, 1, 0), (23, 200, 160, 2, 1, 2), (24, 512, 512, 12, 2, 3)])
def test_mixed_lights_float4(ctx, oracle, seed, w, h, ns, nd, nl):
    s = scenes.lighting_scene(seed, w, h, ns, n_directional=nd, n_line=nl, ramp=(60.0, 220.0), ao=True, float4_lightmap=True)
    r, tex = make_renderer(ctx, s)
    _check(r.RenderLighting(), oracle_lightmap(oracle, r, tex, s), f"mixed {seed}")


def test_half4_and_rgba8_lightmaps(ctx, oracle):
    s = scenes.lighting_scene(25, 256, 192, 5, n_directional=1, ramp=(60.0, 200.0))
    r, tex = make_renderer(ctx, s)
    s.configuration.Float4Lightmap = True
    ref = oracle_lightmap(oracle, r, tex, s)
    s.configuration.Float4Lightmap = False
    half = r.RenderLighting()
    assert half.dtype == np.float16
    want = oracle.float_to_half(ref)
    # identical up to one half ulp where the fp32 values differ by <= 1e-4 relative
    diff = np.abs(half.view(np.uint16).astype(np.int32) - want.view(np.uint16).astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    s.configuration.HighQuality = False
    rgba = r.RenderLighting()
    assert rgba.dtype == np.uint8
    want8 = np.floor(np.clip(ref, 0, 1) * 255 + 0.5).astype(np.int32)
    assert np.abs(rgba.astype(np.int32) - want8).max() <= 1


def test_row_bands_tile_the_frame(ctx, oracle):
    s = scenes.lighting_scene(26, 300, 210, 5, n_directional=1, n_line=1, ramp=(60.0, 200.0), float4_lightmap=True)
    r, tex = make_renderer(ctx, s)
    full = r.RenderLighting()
    bands = [r.RenderLighting(rows=(a, b)) for a, b in ((0, 53), (53, 106), (106, 107), (107, 210))]
    assert np.array_equal(np.concatenate(bands, axis=0), full)   # bit-identical: shards do not change results
    _check(full, oracle_lightmap(oracle, r, tex, s), "bands")


def test_no_distance_field_and_no_gbuffer(ctx, oracle):
    s = scenes.lighting_scene(27, 200, 120, 4, n_directional=1, ramp=(50.0, 120.0), float4_lightmap=True)
    s.configuration.EnableGBuffer = False
    r = ib.LightingRenderer(ctx, s.environment, s.configuration)
    r.SetGBuffer(None)
    lm = r.RenderLighting()
    ref = oracle_lightmap(oracle, r, None, s)
    _check(lm, ref, "no df / no gbuffer")
    # closed form: unobstructed, flat ground => ambient + sum(color * a * falloff); alpha counts lights (AllowFullbright off w/o gbuffer)
    assert lm[..., 3].min() >= s.environment.Ambient[3]


def test_shadow_filter_fullbright_and_unshadowed_pixels(ctx, oracle):
    s = scenes.lighting_scene(28, 192, 128, 3, n_directional=1, ramp=(60.0, 160.0), float4_lightmap=True)
    rs = np.random.RandomState(5)
    h, w = s.gbuffer.shape[:2]
    z = rs.uniform(0, 40, (h, w)).astype(np.float32)
    n = np.zeros((h, w, 3), np.float32); n[..., 2] = 1
    tilt = rs.rand(h, w) < 0.3
    n[tilt] = np.array([0.6, 0.0, 0.8], np.float32)
    n[rs.rand(h, w) < 0.05] = 0          # "no normal" pixels
    es = rs.rand(h, w) > 0.3
    fb = rs.rand(h, w) < 0.1
    dead = rs.rand(h, w) < 0.05
    s.gbuffer = ib.encode_gbuffer(n, np.zeros((h, w), np.float32), z, es, fb, dead)
    s.environment.Lights[0].ShadowFilter = ib.ShadowFilter.Shadowed
    s.environment.Lights[1].ShadowFilter = ib.ShadowFilter.Unshadowed
    s.environment.Lights[2].SpecularColor = (0.4, 0.3, 0.2)
    s.environment.Lights[2].SpecularPower = 6.0
    for stencil in (False, True):
        s.configuration.StencilCulling = stencil
        r, tex = make_renderer(ctx, s)
        _check(r.RenderLighting(), oracle_lightmap(oracle, r, tex, s), f"flags stencil={stencil}")


def test_quality_falloff_modes_and_yfactor(ctx, oracle):
    q = ib.RendererQualitySettings(MinStepSize=1.0, LongStepFactor=0.5, MaxStepCount=64, MaxConeRadius=24, OcclusionToOpacityPower=0.7)
    s = scenes.lighting_scene(29, 256, 160, 4, ramp=(50.0, 140.0), quality=q, float4_lightmap=True)
    L = s.environment.Lights
    L[0].RampMode = ib.LightSourceRampMode.None_
    L[0].Radius = 40.0
    L[1].FalloffYFactor = 2.5
    L[2].FalloffYFactor = 0.6          # the quad clips lit pixels: coverage test matters
    L[3].Quality = ib.RendererQualitySettings(MaxStepCount=8)   # second batch; step-limit ramp
    L[3].ShadowDistanceFalloff = 30.0
    s.configuration.LightOcclusion = 20.0
    r, tex = make_renderer(ctx, s)
    _check(r.RenderLighting(), oracle_lightmap(oracle, r, tex, s), "quality/falloff")


def test_light_probes(ctx, oracle):
    s = scenes.lighting_scene(30, 256, 256, 5, n_directional=1, n_line=2, n_probes=64, ramp=(60.0, 200.0))
    s.probes[3].Normal = None
    s.probes[4].EnableShadows = False
    r, tex = make_renderer(ctx, s)
    gpu = r.UpdateLightProbes(float4=True)
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    pos = np.array([list(p.Position) + [1.0] for p in s.probes], np.float32)
    nrm = np.array([(list(p.Normal) if p.Normal is not None else [0, 0, 0]) + [1.0 if p.EnableShadows else 0.0] for p in s.probes], np.float32)
    ref = oracle.update_light_probes(tex, frame, batches, nb, verts, nv, pos, nrm)
    _check(gpu, ref, "probes")
    half = r.UpdateLightProbes()
    assert half.dtype == np.float16 and half.shape == (64, 4)


def test_render_scale_and_viewport(ctx, oracle):
    s = scenes.lighting_scene(31, 240, 160, 4, n_directional=1, ramp=(60.0, 160.0), float4_lightmap=True)
    s.configuration.RenderScale = (0.5, 0.5)
    s.gbuffer = s.gbuffer[:80, :120].copy()
    r, tex = make_renderer(ctx, s)
    r.ViewportPosition = (10.0, -6.0)
    r.ViewportScale = (1.25, 1.25)
    lm = r.RenderLighting()
    assert lm.shape == (80, 120, 4)
    _check(lm, oracle_lightmap(oracle, r, tex, s), "render scale")


def test_empty_and_degenerate_inputs(ctx):
    s = scenes.lighting_scene(32, 64, 48, 0, float4_lightmap=True)
    r, _ = make_renderer(ctx, s)
    lm = r.RenderLighting()
    amb = np.array(s.environment.Ambient, np.float32); amb[3] = 0   # fullbright mode zeroes alpha
    assert np.array_equal(lm, np.broadcast_to(amb, lm.shape))
    assert r.RenderLighting(rows=(10, 10)).shape == (0, 64, 4)
    s.environment.Lights = [ib.SphereLightSource(Position=(10, 10, 5), Radius=4, RampLength=20, Opacity=0.0)]
    assert np.array_equal(r.RenderLighting(), np.broadcast_to(amb, lm.shape))   # Opacity <= 0 lights are skipped on the host


def test_error_reporting(ctx):
    import ctypes as C
    from illuminant_b200 import _abi
    s = scenes.lighting_scene(33, 64, 48, 1, float4_lightmap=True)
    r, _ = make_renderer(ctx, s)
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    batches[0].light_type = 5   # Projector: out of scope
    out = np.empty((48, 64, 4), np.float32)
    rc = ctx.lib.ilb_render_lighting(ctx.handle, r.DistanceField.handle, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                     C.cast(verts, C.c_void_p), nv, out.ctypes.data_as(C.c_void_p))
    assert rc == _abi.ERR_UNSUPPORTED and b"light type" in ctx.lib.ilb_last_error(ctx.handle)
    with pytest.raises(ib.IlluminantError):
        r.DistanceField.Load(np.zeros(16, np.uint16))   # "Truncated file"


def test_planes_match_atlas_bit_for_bit(ctx, monkeypatch):
    """The expanded-planes sampler (csrc/planes.cu, sampleFieldPlanesT) and the Rgba64-atlas sampler are two layouts of
    the same arithmetic: lightmaps and probes must be bit-identical, including a 3-column atlas (1/3 is inexact),
    a reduced-resolution field and rays that leave the volume."""
    for seed, w, h, slices, res in ((31, 300, 210, 8, 1.0), (32, 200, 160, 24, 0.5), (33, 257, 131, 3, 1.0)):
        s = scenes.lighting_scene(seed, w, h, 6, n_directional=2, n_line=2, n_probes=16, ramp=(60.0, 260.0), ao=True, float4_lightmap=True)
        s.df_slices = slices
        out = []
        for flag in ("1", "0"):
            monkeypatch.setenv("ILB_NO_PLANES", flag)
            df = scenes.make_distance_field(ctx, s, resolution=res)
            df.Rasterize(s.obstructions)
            r = ib.LightingRenderer(ctx, s.environment, s.configuration)
            r.DistanceField = df
            r.Probes = s.probes
            r.SetGBuffer(s.gbuffer)
            out.append((r.RenderLighting(), r.UpdateLightProbes() if s.probes else None))
        assert np.array_equal(out[0][0], out[1][0]), f"seed {seed}"
        if out[0][1] is not None:
            assert np.array_equal(out[0][1], out[1][1])


def test_pipelined_host_frame_equals_upload_plus_render(ctx):
    """ilb_render_lighting_frame (G-buffer up / shade / lightmap down, pipelined over row bands on three streams) returns
    the same bits as ilb_gbuffer_upload + ilb_render_lighting, for whole frames, bands and odd sizes."""
    for seed, w, h in ((34, 300, 211), (35, 640, 400)):
        s = scenes.lighting_scene(seed, w, h, 6, n_directional=1, n_line=1, ramp=(60.0, 260.0))
        df = scenes.make_distance_field(ctx, s)
        df.Rasterize(s.obstructions)
        r = ib.LightingRenderer(ctx, s.environment, s.configuration)
        r.DistanceField = df
        r.SetGBuffer(s.gbuffer)
        for rows in (None, (16, h - 37)):
            want = r.RenderLighting(rows=rows)
            r.SetGBuffer(np.zeros_like(s.gbuffer))          # the pipelined call must bring its own G-buffer
            got = r.RenderLightingFrame(s.gbuffer, rows=rows)
            assert np.array_equal(got.view(np.uint16), want.view(np.uint16))


def test_bands_land_in_one_registered_shared_host_frame(ctx):
    """The host-to-host leg at N > 1 on one device: the bands of a frame, each rendered by its own ilb_render_lighting_frame
    call straight into its rows of ONE page-locked shared-memory host frame (ilb_host_register), reassemble the unsharded
    lightmap bit for bit -- under every split of the pipeline bands; registering twice and unregistering unknown memory are
    harmless; null ranges are rejected."""
    import os
    from illuminant_b200 import sharding
    from illuminant_b200._abi import IlluminantError
    w, h = 640, 403
    s = scenes.lighting_scene(37, w, h, 6, n_directional=1, n_line=1, ramp=(60.0, 260.0))
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    r = ib.LightingRenderer(ctx, s.environment, s.configuration)
    r.DistanceField = df
    r.SetGBuffer(s.gbuffer)
    want = r.RenderLighting()
    shared = sharding.SharedHostFrame(f"ilb_test_gpu_frame_{os.getpid()}", h, w, 4, want.dtype, 0, 1, ctx=ctx)
    try:
        ctx.host_register(shared.frame.ctypes.data - shared.HEADER, shared.HEADER + shared.frame.nbytes)   # already registered: fine
        gb = np.ascontiguousarray(s.gbuffer)
        old = os.environ.get("ILB_BAND_SHARES")
        try:
            for split in (None, "0", "1", "2", "3", "4"):
                if split is None:
                    os.environ.pop("ILB_BAND_SHARES", None)
                else:
                    os.environ["ILB_BAND_SHARES"] = split
                shared.frame[...] = 0
                r.SetGBuffer(np.zeros_like(s.gbuffer))
                for k in range(3):
                    r0, r1 = sharding.row_band(k, 3, h)
                    r.RenderLightingFrame(gb, rows=(r0, r1), out=shared.rows(r0, r1))
                assert np.array_equal(shared.frame.view(np.uint16), want.view(np.uint16)), split
        finally:
            if old is None:
                os.environ.pop("ILB_BAND_SHARES", None)
            else:
                os.environ["ILB_BAND_SHARES"] = old
        with pytest.raises(IlluminantError):
            ctx.host_register(0, 4096)
        scratch = np.zeros(8192, np.uint8)
        ctx.host_unregister(scratch.ctypes.data)      # never registered: nothing to do
    finally:
        shared.close()


def test_frames_in_flight_are_chained_band_by_band(ctx):
    """ilb_render_lighting_frame_async / _wait: several frames queued before the first is waited for -- every frame with its own
    G-buffer and its own light list, into its own page-locked output -- give the bits of the synchronous call; other entry points
    in between (a G-buffer upload, a plain render, a frame of another size) first wait for what is in flight; tickets of frames
    that are long complete, of empty bands and unknown tickets behave as documented."""
    from illuminant_b200._abi import IlluminantError
    w, h = 640, 403
    rng = np.random.RandomState(5)
    base = scenes.lighting_scene(51, w, h, 6, n_directional=1, n_line=1, ramp=(60.0, 260.0))
    df = scenes.make_distance_field(ctx, base)
    df.Rasterize(base.obstructions)
    r = ib.LightingRenderer(ctx, base.environment, base.configuration)
    r.DistanceField = df
    gbs, wants = [], []
    sphere = next(l for l in base.environment.Lights if isinstance(l, ib.SphereLightSource))
    x0 = sphere.Position[0]
    n_frames = 6
    for k in range(n_frames):      # frame k: its own G-buffer (rows shuffled in blocks) and a moved light
        gb = np.ascontiguousarray(np.roll(base.gbuffer, 16 * k, axis=0), dtype=np.float32 if base.configuration.HighQualityGBuffer else np.float16)
        sphere.Position = (x0 + 7.0 * k,) + tuple(sphere.Position[1:])
        gbs.append(gb)
        wants.append(r.RenderLightingFrame(gb).copy())
    pinned = [np.zeros_like(wants[0]) for _ in range(n_frames)]
    for a in gbs + pinned:
        ctx.host_register(a.ctypes.data, a.nbytes)
    try:
        for rows in (None, (32, h - 40)):
            for o in pinned:
                o[...] = 0
            tickets = []
            for k in range(n_frames):      # all queued before the first wait: up to six frames in flight
                sphere.Position = (x0 + 7.0 * k,) + tuple(sphere.Position[1:])
                out = pinned[k] if rows is None else pinned[k][rows[0]:rows[1]]
                tickets.append(r.RenderLightingFrameAsync(gbs[k], out, rows=rows))
            assert tickets == sorted(tickets) and len(set(tickets)) == n_frames
            for k in reversed(range(n_frames)):    # waiting out of order is allowed
                r.WaitLightingFrame(tickets[k])
                want = wants[k] if rows is None else wants[k][rows[0]:rows[1]]
                got = pinned[k] if rows is None else pinned[k][rows[0]:rows[1]]
                assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), (rows, k)
            r.WaitLightingFrame(tickets[0])        # long complete: returns at once
        # other entry points between frames in flight
        sphere.Position = (x0,) + tuple(sphere.Position[1:])
        t0 = r.RenderLightingFrameAsync(gbs[0], pinned[0])
        r.SetGBuffer(gbs[1])                       # waits for t0, then replaces the G-buffer
        plain = r.RenderLighting()                 # renders from gbs[1]
        t1 = r.RenderLightingFrameAsync(gbs[2], pinned[2])
        t2 = r.RenderLightingFrameAsync(gbs[3], pinned[3], rows=(0, 160))     # another geometry: not chained
        r.WaitLightingFrame(t2)
        r.WaitLightingFrame(t1)
        r.WaitLightingFrame(t0)
        sphere.Position = (x0,) + tuple(sphere.Position[1:])
        for k, got in ((0, pinned[0]), (2, pinned[2])):
            assert np.array_equal(got.view(np.uint16), r.RenderLightingFrame(gbs[k]).view(np.uint16)), k
        assert np.array_equal(pinned[3][:160].view(np.uint16), r.RenderLightingFrame(gbs[3], rows=(0, 160)).view(np.uint16))
        assert np.array_equal(plain.view(np.uint16), r.RenderLightingFrame(gbs[1]).view(np.uint16))
        assert r.RenderLightingFrameAsync(gbs[0], pinned[0][:0], rows=(48, 48)) == 0      # empty band: nothing queued
        r.WaitLightingFrame(0)
        with pytest.raises(IlluminantError):
            r.WaitLightingFrame(10 ** 9)
        with pytest.raises(ValueError):
            r.RenderLightingFrameAsync(gbs[0][:, ::2], pinned[0])
    finally:
        ctx.synchronize()
        for a in gbs + pinned:
            ctx.host_unregister(a.ctypes.data)


def test_heaviest_first_tile_order_is_only_a_schedule(ctx):
    """ILB_OPT_LIGHT_TILE_ORDER starts the tiles under the most sphere-light quads first: same bits with the option on and off,
    for whole frames, bands and mixed / sphere-only light lists; moving a light makes a new order (more moves than the cache
    has slots: slots are recycled while frames are in flight)."""
    from illuminant_b200 import _abi
    w, h = 800, 480      # 50 x 30 tiles: above the 512-tile threshold
    for n_line in (1, 0):
        s = scenes.lighting_scene(41 + n_line, w, h, 10, n_directional=1, n_line=n_line, ramp=(40.0, 160.0))
        df = scenes.make_distance_field(ctx, s)
        df.Rasterize(s.obstructions)
        r = ib.LightingRenderer(ctx, s.environment, s.configuration)
        r.DistanceField = df
        r.SetGBuffer(s.gbuffer)
        try:
            for rows in (None, (16, h - 21)):
                ctx.set_option(_abi.OPT_LIGHT_TILE_ORDER, 0)
                want = r.RenderLighting(rows=rows)
                ctx.set_option(_abi.OPT_LIGHT_TILE_ORDER, 1)
                for _ in range(2):     # miss, then hit
                    got = r.RenderLighting(rows=rows)
                    assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), (n_line, rows)
            light = next(l for l in s.environment.Lights if isinstance(l, ib.SphereLightSource))
            x0 = light.Position[0]
            for k in range(40):
                light.Position = (x0 + 3.0 * k,) + tuple(light.Position[1:])
                ctx.set_option(_abi.OPT_LIGHT_TILE_ORDER, 1)
                got = r.RenderLighting()
                if k % 13 == 0:
                    ctx.set_option(_abi.OPT_LIGHT_TILE_ORDER, 0)
                    assert np.array_equal(got.view(np.uint16), r.RenderLighting().view(np.uint16)), k
        finally:
            ctx.set_option(_abi.OPT_LIGHT_TILE_ORDER, 0)


def test_degenerate_geometry_takes_the_ieee_fallback(ctx, oracle):
    """Operands outside the fast window of the deferred-guard square roots / reciprocals (zero-length vectors: a light
    exactly at a shaded point, at a trace origin, a pixel on a line light's axis, a zero-length line light) must come out
    of the IEEE re-evaluation exactly like the oracle, NaN for NaN."""
    s = scenes.lighting_scene(36, 160, 96, 0, float4_lightmap=True)
    s.gbuffer[...] = scenes.make_gbuffer(np.random.RandomState(1), 160, 96, 0)       # flat ground: world = (px + 0.5, py + 0.5, 0)
    c = (0.8, 0.6, 0.4, 1.0)
    s.environment.Lights = [
        ib.SphereLightSource(Position=(40.5, 30.5, 0.0), Radius=6.0, RampLength=60.0, Color=c, CastsShadows=True),     # on a shaded point
        ib.SphereLightSource(Position=(80.5, 50.5, 1.6), Radius=4.0, RampLength=50.0, Color=c, CastsShadows=True),     # on a trace origin
        ib.SphereLightSource(Position=(20.5, 70.5, 0.0), Radius=0.0, RampLength=40.0, Color=c, CastsShadows=False),
        ib.LineLightSource(StartPosition=(100.5, 20.5, 0.0), EndPosition=(140.5, 20.5, 0.0), Radius=5.0, StartColor=c, EndColor=c, CastsShadows=True),
        ib.LineLightSource(StartPosition=(60.5, 80.5, 12.0), EndPosition=(60.5, 80.5, 12.0), Radius=5.0, StartColor=c, EndColor=c, CastsShadows=True),
        ib.LineLightSource(StartPosition=(10.5, 10.5, 1.5), EndPosition=(10.5, 40.5, 1.5), Radius=3.0, StartColor=c, EndColor=c, CastsShadows=True),
    ]
    r, tex = make_renderer(ctx, s)
    gpu, ref = r.RenderLighting(), oracle_lightmap(oracle, r, tex, s)
    assert np.array_equal(np.isnan(gpu), np.isnan(ref))
    ok = ~np.isnan(ref)
    err = lighting_rel_err(gpu[ok], ref[ok])
    assert err.max() <= LIGHTING_RTOL, f"max rel err {err.max():.3e}"


def test_dynamic_distance_field(ctx, oracle):
    """DynamicDistanceField (SDF/DistanceField.cs:248-310): sampled field = static field with the dynamic obstructions
    MAX-blended on top, rewritten in place per frame; the derived planes follow the new contents."""
    s = scenes.lighting_scene(37, 320, 200, 5, n_directional=1, ramp=(60.0, 220.0), float4_lightmap=True)
    rs = np.random.RandomState(5)
    dynamic = [ib.LightObstruction(ib.LightObstructionType.Ellipsoid, (float(rs.uniform(40, 280)), float(rs.uniform(30, 170)), 10.0), (18.0, 12.0, 40.0),
                                   IsDynamic=True) for _ in range(3)]
    df = ib.DynamicDistanceField(ctx, s.width, s.height, s.df_depth, s.df_slices, s.df_resolution, 128)
    df.Rasterize(s.obstructions + dynamic)
    static_ref = oracle.generate_distance_field(df, s.obstructions)
    assert np.abs(df.SaveStatic().astype(np.int32) - static_ref.astype(np.int32)).max() <= 1
    static_gpu = df.SaveStatic()
    ref = oracle.generate_distance_field(df, dynamic, base=static_gpu)
    got = df.Save()
    assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1 and (got != ref).mean() < 1e-3
    # all obstructions at once is the same field (MAX is associative; re-quantising the static part is idempotent)
    whole = scenes.make_distance_field(ctx, s)
    whole.Rasterize(s.obstructions + dynamic)
    assert np.array_equal(whole.Save(), got)

    r = ib.LightingRenderer(ctx, s.environment, s.configuration)
    r.DistanceField = df
    r.SetGBuffer(s.gbuffer)
    before = r.RenderLighting()
    _check(before, oracle_lightmap(oracle, r, got, s), "dynamic field, frame 0")
    for o in dynamic:                                   # next frame: the dynamic obstructions moved
        o.Center = (o.Center[0] + 25.0, o.Center[1] - 10.0, o.Center[2])
    df.RasterizeDynamic(dynamic)
    moved = df.Save()
    assert not np.array_equal(moved, got)
    after = r.RenderLighting()                          # same handle, rewritten atlas: the planes must have been refreshed
    assert not np.array_equal(after, before)
    _check(after, oracle_lightmap(oracle, r, moved, s), "dynamic field, frame 1")


def test_particle_light_source(ctx, oracle):
    """ParticleLightSource ("next" row N4; ParticleLight.fx, LightingRenderer.cs:769-789, :1126-1144): every live particle
    with a visible colour lights the frame like a sphere light with the template's properties; the light list is built on
    the device from the particle state.  The oracle gets the same lights as LightVertex records."""
    s = scenes.lighting_scene(38, 320, 200, 2, n_directional=1, n_line=1, ramp=(60.0, 160.0), float4_lightmap=True)
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    tex = df.Save()
    rs = np.random.RandomState(9)
    n = 700
    pos = np.zeros((n, 4), np.float32)
    pos[:, 0], pos[:, 1], pos[:, 2] = rs.uniform(0, 320, n), rs.uniform(0, 200, n), rs.uniform(2, 40, n)
    pos[:, 3] = rs.uniform(-0.5, 3.0, n)                       # some are dead
    vel = np.zeros((n, 4), np.float32)
    attr = rs.uniform(0.0, 1.0, (n, 4)).astype(np.float32)
    attr[::7, 3] = 0.0                                         # some are invisible
    attr[:, :3] *= attr[:, 3:4]                                # premultiplied like the spawner output
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=16, RandomSeed=1))
    system = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=4)
    system.Spawn(pos, vel, attr)
    template = ib.SphereLightSource(Radius=3.0, RampLength=24.0, Color=(0.9, 0.7, 0.5, 0.6), SpecularColor=(0.2, 0.2, 0.3), SpecularPower=4.0,
                                    CastsShadows=True, AmbientOcclusionRadius=6.0, AmbientOcclusionOpacity=0.5)
    pls = ib.ParticleLightSource(Template=template, System=system)
    s.environment.Lights.append(pls)
    r = ib.LightingRenderer(ctx, s.environment, s.configuration)
    r.DistanceField = df
    r.SetGBuffer(s.gbuffer)
    gpu = r.RenderLighting()

    # oracle: the host lights' batches plus one ILB_LIGHT_PARTICLE batch with the vertex-shader outputs, in particle order
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    per = 16 * 16
    P = np.zeros((system.LiveChunkCount * per, 4), np.float32)
    A = np.zeros_like(P)
    P[:n], A[:n] = pos, attr
    pv = pls.light_vertices(P, A, True)
    assert 300 < len(pv) < n
    from illuminant_b200._abi import LightBatch, LightVertex
    allv = (LightVertex * (nv + len(pv)))(*([verts[i] for i in range(nv)] + pv))
    allb = (LightBatch * (nb + 1))(*[batches[i] for i in range(nb)])
    allb[nb].light_type, allb[nb].first_vertex, allb[nb].vertex_count = 3, nv, len(pv)
    allb[nb].df = r._df_uniforms(None)
    ref = oracle.render_lighting(tex, s.gbuffer, frame, allb, nb + 1, allv, nv + len(pv))
    _check(gpu, ref, "particle lights")
    assert np.array_equal(gpu[..., 3], ref[..., 3])            # alpha counts the lights touching each pixel
    # the source really contributes, and goes away with the list
    pls.Enabled = False
    assert not np.array_equal(r.RenderLighting(), gpu)
    bands = None
    pls.Enabled = True
    bands = [r.RenderLighting(rows=(a, b)) for a, b in ((0, 64), (64, 200))]
    assert np.array_equal(np.concatenate(bands, axis=0), gpu)


def test_light_list_caching_and_the_shared_constant_bank(ctx):
    """The flattened light records are uploaded only when they change, and frames of up to 256 lights read them from the
    constant bank -- which belongs to the module, i.e. to every context of the process.  Two contexts that render different
    light lists alternately, a light that moves between two frames, and the option switched off must all give the same bits
    as a fresh render."""
    from illuminant_b200 import _abi
    sa = scenes.lighting_scene(51, 160, 96, 5, n_directional=1, n_line=1, ramp=(50.0, 140.0), float4_lightmap=True)
    sb = scenes.lighting_scene(52, 160, 96, 3, n_line=2, ramp=(60.0, 120.0), float4_lightmap=True)
    ra, _ = make_renderer(ctx, sa)
    c2 = ib.Context(0)
    try:
        rb, _ = make_renderer(c2, sb)
        a0, b0 = ra.RenderLighting(), rb.RenderLighting()
        for _ in range(3):      # unchanged lists: nothing is uploaded, and each context finds the other's records in the bank
            assert np.array_equal(ra.RenderLighting(), a0) and np.array_equal(rb.RenderLighting(), b0)
        ctx.set_option(_abi.OPT_LIGHT_CONST_BANK, 0)
        assert np.array_equal(ra.RenderLighting(), a0)          # the global-memory path: same bits
        ctx.set_option(_abi.OPT_LIGHT_CONST_BANK, 1)
        assert np.array_equal(ra.RenderLighting(), a0)
        light = sa.environment.Lights[0]
        x, y, z = light.Position
        light.Position = (x + 7.0, y - 3.0, z)
        a1 = ra.RenderLighting()
        assert not np.array_equal(a1, a0)                        # the moved light was uploaded
        light.Position = (x, y, z)
        assert np.array_equal(ra.RenderLighting(), a0) and np.array_equal(rb.RenderLighting(), b0)
    finally:
        c2.close()


def test_line_lights_on_tilted_and_flat_surfaces(ctx, oracle):
    """The line light's opacity has a short form for surfaces whose normal has no x / y component (floors, box tops) and the
    general form for everything else: a G-buffer that mixes both inside every warp, normals that point away from the light,
    and missing normals must all match the oracle (which only has the general form), alpha counts exactly."""
    s = scenes.lighting_scene(53, 224, 144, 2, n_directional=1, n_line=4, ramp=(60.0, 160.0), float4_lightmap=True)
    rs = np.random.RandomState(12)
    h, w = s.gbuffer.shape[:2]
    z = rs.uniform(0, 30, (h, w)).astype(np.float32)
    n = np.zeros((h, w, 3), np.float32)
    n[..., 2] = 1
    tilt = rs.rand(h, w) < 0.5
    v = rs.normal(size=(h, w, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    n[tilt] = v[tilt]                       # any direction, also facing down
    n[rs.rand(h, w) < 0.1] = np.array([0, 0, -1], np.float32)
    n[rs.rand(h, w) < 0.05] = 0             # "no normal" pixels
    s.gbuffer = ib.encode_gbuffer(n, np.zeros((h, w), np.float32), z)
    r, tex = make_renderer(ctx, s)
    gpu, ref = r.RenderLighting(), oracle_lightmap(oracle, r, tex, s)
    _check(gpu, ref, "line lights, mixed normals")
    assert np.array_equal(gpu[..., 3], ref[..., 3])
