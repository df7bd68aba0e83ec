"""The deliberate conventions of this implementation where the reference leaves behaviour to D3D / the shader compiler
("Conventions ... documented, not chased"), each pinned by a test that asserts the chosen behaviour on the
CPU oracle (the CUDA path is compared with the oracle everywhere else), plus the cross-check kit that lets a machine with the
real reference test the oracle itself."""
import subprocess
import sys
from pathlib import Path

import numpy as np

import illuminant_b200 as ib
from illuminant_b200 import scenes

ROOT = Path(__file__).resolve().parent.parent


def _render(oracle, lights, w=40, h=24):
    s = scenes.lighting_scene(0, w, h, 0, float4_lightmap=True)
    s.configuration.EnableGBuffer = False
    s.environment.Lights = lights
    df = scenes.make_distance_field(None, s)
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField = df
    batches, nb, verts, nv = r.build_batches()
    return oracle.render_lighting(tex, None, r.build_frame(), batches, nb, verts, nv), s


def test_normalize_of_the_zero_vector_is_zero_not_nan(oracle):
    """ps_3_0 `nrm` multiplies by rsq(dot), and rsq(0) is the largest float, so normalize(0) = 0; IEEE 0 * inf would be NaN.
    A line light whose end points coincide makes lightLeft = normalize(P1 - P0) the zero vector (FBPBR.fxh:53-60)."""
    degenerate = ib.LineLightSource(StartPosition=(20.0, 12.0, 10.0), EndPosition=(20.0, 12.0, 10.0), Radius=6.0,
                                    StartColor=(1.0, 1.0, 1.0, 1.0), EndColor=(1.0, 1.0, 1.0, 1.0), CastsShadows=False)
    lm, s = _render(oracle, [degenerate])
    assert np.isfinite(lm).all()
    assert (lm[..., :3] >= np.array(s.environment.Ambient[:3]) - 1e-6).all()      # a light never darkens


def test_specular_term_is_skipped_for_a_black_specular_colour(oracle):
    """CalcSphereLightSpecularity (LightCommon.fxh:212-222) is multiplied by the specular colour; with the default (0, 0, 0) the
    implementation does not evaluate it, so SpecularPower = 0 (pow(0, 0)) cannot inject a NaN."""
    base = dict(Position=(20.0, 12.0, 8.0), Radius=4.0, RampLength=30.0, Color=(0.9, 0.8, 0.7, 1.0), CastsShadows=False)
    plain, _ = _render(oracle, [ib.SphereLightSource(**base)])
    power0, _ = _render(oracle, [ib.SphereLightSource(SpecularColor=(0.0, 0.0, 0.0), SpecularPower=0.0, **base)])
    assert np.isfinite(power0).all() and np.array_equal(plain, power0)
    lit, _ = _render(oracle, [ib.SphereLightSource(SpecularColor=(0.5, 0.5, 0.5), SpecularPower=8.0, **base)])
    assert (lit[..., :3] >= plain[..., :3] - 1e-7).all() and (lit[..., :3] > plain[..., :3] + 1e-4).any()   # and it works when asked for


def test_tone_mapped_resolve_clamps_negative_light_at_zero(oracle):
    """Resolve.fx:117-140: the tone-mapping curve is evaluated on max(light, 0); a negative lightmap texel (possible with negative
    light colours) resolves to black instead of running the rational curve outside its domain."""
    from illuminant_b200 import _abi, hdr
    cfg = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.0, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=3.0))
    p = hdr.pack_resolve(2, 1, _abi.FORMAT_FLOAT4, cfg, _abi.FORMAT_FLOAT4, _abi.FORMAT_FLOAT4)
    lm = np.zeros((1, 2, 4), np.float32)
    lm[0, 0, :3], lm[0, 1, :3] = -0.75, 0.5
    out = oracle.resolve_lighting(p, lm)
    assert np.isfinite(out).all() and np.allclose(out[0, 0, :3], 0.0, atol=1e-6) and (out[0, 1, :3] > 0.05).all()


def test_lightmap_is_summed_in_fp32_and_rounded_once(oracle):
    """The reference's HalfVector4 target rounds after every light's additive pass; here the sum stays in fp32 and is rounded to
    half once per pixel.  The half4 output is therefore exactly the rounding of the fp32 lightmap (checked on the GPU at full size
    in test_c2_1080p_32_sphere_lights_full_frame_parity); here: many dim lights whose individual contributions are below half
    precision at the running sum still add up."""
    n = 64
    lights = [ib.SphereLightSource(Position=(20.0, 12.0, 6.0), Radius=40.0, RampLength=10.0, Color=(1.0, 1.0, 1.0, 2.0e-4), CastsShadows=False)
              for _ in range(n)]
    lm, s = _render(oracle, lights)
    ambient = np.float32(s.environment.Ambient[0])
    got = lm[12, 20, 0]
    assert abs(got - (ambient + n * 2.0e-4)) < 1e-6
    per_light_half = np.float16(ambient)
    for _ in range(n):                                   # what rounding after every pass would give
        per_light_half = np.float16(np.float32(per_light_half) + np.float32(2.0e-4))
    assert abs(float(np.float16(got)) - float(got)) <= 4e-5 < abs(float(per_light_half) - float(got))


def test_crosscheck_kit_is_generated(tmp_path):
    """tools/crosscheck: field in DistanceField.Save layout + scene + oracle lightmap + the TestGame scene source."""
    out = tmp_path / "kit"
    res = subprocess.run([sys.executable, str(ROOT / "tools" / "crosscheck" / "make_kit.py"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    field, lm = (out / "field.rgba64").stat().st_size, (out / "oracle_lightmap.f32").stat().st_size
    assert lm == 512 * 384 * 16 and field % 8 == 0 and field >= 512 * 384 * 8 * 3
    cs = (out / "CrossCheckScene.cs").read_text()
    assert "DistanceField.Load(stream)" in cs and cs.count("Environment.Lights.Add(") == 9 and "reference_lightmap.f32" in cs
    cmp_ = subprocess.run([sys.executable, str(ROOT / "tools" / "crosscheck" / "compare.py"), str(out / "oracle_lightmap.f32"),
                           str(out / "oracle_lightmap.f32"), "512", "384"], capture_output=True, text=True)
    assert cmp_.returncode == 0 and "max relative error 0.000e+00" in cmp_.stdout
