"""illuminant_b200 -- B200-native (sm_100a) implementation of sq/Illuminant's two data-parallel hot paths:
the LightingRenderer SDF cone-trace and the ParticleEngine update chain, behind the C-ABI of
include/illuminant_b200.h.  This package is the host-side mirror of the reference API for those paths."""
from ._abi import (Context, IlluminantError, EXPORTED_SYMBOLS, LIB_PATH, load_library,
                   FORMAT_FLOAT4, FORMAT_HALF4, FORMAT_RGBA8)
from .distance_field import (DistanceField, DynamicDistanceField, LightObstruction, LightObstructionType, RendererQualitySettings,
                             SimpleHeightVolume)
from .lighting import (DirectionalLightSource, LightingEnvironment, LightingRenderer, LightProbe, LightSourceRampMode,
                       LineLightSource, ParticleLightSource, RendererConfiguration, ShadowFilter, SphereLightSource, encode_gbuffer)
from .hdr import (ColorLUT, DitheringSettings, GammaCompressionConfiguration, HDRConfiguration, HDRMode, Histogram, LUTBlendingConfiguration,
                  RenderedLighting, ToneMappingConfiguration, pack_resolve)
from .particles import (FMA, AreaType, Attractor, AttractorType, Bezier4V, BezierF, FeedbackSpawner, Formula, FormulaType, Gravity, MatrixMultiply,
                        Noise, ParticleAppearance, ParticleCollision, ParticleColorLifeRamp, ParticleEngine, ParticleEngineConfiguration, ParticleRenderParameters, ParticleSystem,
                        ParticleSystemConfiguration, PatternSpawner, Spawner, TransformArea)

__version__ = "0.1.0"
