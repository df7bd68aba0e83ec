"""Host-side mirror of the reference particle API for the hot path: ParticleEngine, ParticleSystem and the
Spawner / Gravity / Noise / FMA / MatrixMultiply transforms (Illuminant/Particles/*.cs, cited per member).

`ParticleSystem.Update` keeps the reference's order -- spawners first (ParticleSystem.cs:725-741), then for every
chunk the active transforms in list order and the final Update pass (:791-856) -- but issues it as one call into
the C-ABI (`ilb_particles_step`).  `Parameter<T>` values are plain numbers here: they cross the boundary already
evaluated.  Host randomness (spawn counts, RandomnessOffset, Noise U/V) comes from a seedable numpy Generator instead
of the un-vendored CoreCLR.Xoshiro; every draw is passed explicitly, so results depend only on the seed.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _abi
from ._abi import (OP_FMA, OP_GRAVITY, OP_MATRIX_MULTIPLY, OP_NOISE, Area, Bezier1, Bezier4, Float4, Op, PsysUniforms, Spawn, SpawnSource)
from .distance_field import DistanceField

F = np.float32
RandomnessTextureWidth, RandomnessTextureHeight = 807, 653  # ParticleEngine.cs:45-46
VelocityConstantScale = 1000                                # Uniforms.cs:199
MaxInlinePositions = 4                                      # ParticleSpawner.cs:263
IDENTITY = (1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0)


class AreaType:  # TransformArea types == LightObstructionType ids (DistanceFunctionCommon.fxh:167-186)
    None_, Ellipsoid, Box, Cylinder, Spheroid, Octagon = 0, 1, 2, 3, 4, 5


class FormulaType:  # SpawnerCommon.fxh:34-37
    Linear, Spherical, Towards, Rectangular = 0, 1, 2, 3


class AttractorType:  # Transforms.cs:297-301
    Physical, Linear, Exponential = 0, 1, 2


@dataclass
class ParticleEngineConfiguration:  # ParticleEngine.cs:616-696
    ChunkSize: int = 256
    UpdatesPerSecond: Optional[float] = None
    MaximumUpdateDeltaTimeSeconds: float = 1 / 20.0
    RandomSeed: Optional[int] = None


@dataclass
class ParticleCollision:  # ParticleConfiguration.cs:13-45
    DistanceField: Optional[DistanceField] = None
    DistanceFieldMaximumZ: Optional[float] = None
    EscapeVelocity: float = 128.0
    BounceVelocityMultiplier: float = 0.0
    Distance: float = 0.33
    LifePenalty: float = 0.0
    FullFieldAddressing: bool = False   # not in the reference, see collision_field_uniforms


@dataclass
class BezierF:  # Bezier.cs (BezierF feeding ClampedBezier1, :434-459)
    Count: int = 1
    Mode: int = 0
    MinValue: float = 0.0
    MaxValue: float = 1.0
    A: float = 1.0
    B: float = 1.0
    C: float = 1.0
    D: float = 1.0


@dataclass
class Bezier4V:  # Bezier4 feeding ClampedBezier4 (Bezier.cs:589-600)
    Count: int = 1
    Mode: int = 0
    MinValue: float = 0.0
    MaxValue: float = 1.0
    A: tuple = (1.0, 1.0, 1.0, 1.0)
    B: tuple = (1.0, 1.0, 1.0, 1.0)
    C: tuple = (1.0, 1.0, 1.0, 1.0)
    D: tuple = (1.0, 1.0, 1.0, 1.0)


def _range_and_count(src) -> Float4:
    rng = F(src.MaxValue) - F(src.MinValue)
    if rng == 0 or src.Count <= 1:
        rng = F(1)
    return Float4(min(src.MinValue, src.MaxValue), F(1.0) / rng, src.Count, int(src.Mode))


def clamped_bezier1(src: Optional[BezierF]) -> Bezier1:
    b = Bezier1()
    if src is None:  # ClampedBezier1.One
        b.RangeAndCount, b.ABCD = Float4(0, 1, 1, 0), Float4(1, 1, 1, 1)
        return b
    b.RangeAndCount = _range_and_count(src)
    b.ABCD = Float4(src.A, src.B, src.C, src.D)
    return b


def clamped_bezier4(src: Optional[Bezier4V]) -> Bezier4:
    b = Bezier4()
    if src is None:  # ClampedBezier4.One
        b.RangeAndCount = Float4(0, 1, 1, 0)
        b.A = b.B = b.C = b.D = Float4(1, 1, 1, 1)
        return b
    b.RangeAndCount = _range_and_count(src)
    b.A, b.B, b.C, b.D = Float4(*src.A), Float4(*src.B), Float4(*src.C), Float4(*src.D)
    return b


@dataclass
class ParticleAppearance:  # ParticleConfiguration.cs:42-109
    Texture: Optional[np.ndarray] = None            # uint8 [H, W, 4] (SurfaceFormat.Color) sprite sheet; None = solid quads
    OffsetPx: Tuple[float, float] = (0.0, 0.0)
    SizePx: Optional[Tuple[float, float]] = None
    AnimationRate: Tuple[float, float] = (0.0, 0.0)
    Rounded: bool = False
    DitheredOpacity: bool = False
    RoundingPowerFromLife: Optional["BezierF"] = None   # default BezierF(0.8) (:82)
    Bilinear: bool = True
    RelativeSize: bool = True
    RowFromVelocity: bool = False
    ColumnFromVelocity: bool = False


@dataclass
class ParticleRenderParameters:  # ParticleConfiguration.cs:305-310
    Origin: Tuple[float, float] = (0.0, 0.0)
    Scale: Tuple[float, float] = (1.0, 1.0)
    StippleFactor: Optional[float] = None


@dataclass
class ParticleSystemConfiguration:  # ParticleConfiguration.cs:187-303
    Size: Tuple[float, float] = (1.0, 1.0)
    Friction: float = 0.0
    MaximumVelocity: float = 9999.0
    LifeDecayPerSecond: float = 1.0
    Collision: Optional[ParticleCollision] = None
    RotationFromLife: float = 0.0       # degrees
    RotationFromIndex: float = 0.0      # degrees
    RotationFromVelocity: bool = False
    AnimationRate: Tuple[float, float] = (0.0, 0.0)
    ZToY: float = 0.0
    OpacityFromLife: Optional[float] = None
    ColorFromLife: Optional[Bezier4V] = None
    ColorFromVelocity: Optional[Bezier4V] = None
    SizeFromLife: Optional[BezierF] = None
    SizeFromVelocity: Optional[BezierF] = None
    LifeRamp: Optional["ParticleColorLifeRamp"] = None   # Configuration.Color.LifeRamp (ParticleConfiguration.cs:111-137)
    Appearance: ParticleAppearance = field(default_factory=ParticleAppearance)
    GlobalColor: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)   # Configuration.Color.Global (un-premultiplied)
    SizeFromZ: float = 0.0
    ZFormula: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 0.0)
    StippleFactor: float = 1.0
    WriteRenderOutputs: bool = True     # not in the reference: False skips renderColor/renderData (64 B/particle mode)


@dataclass
class ParticleColorLifeRamp:  # ParticleConfiguration.cs:111-137
    Minimum: float = 0.0
    Maximum: float = 100.0
    Strength: float = 1.0
    Invert: bool = False
    Texture: Optional[np.ndarray] = None   # float32 [H, W, 4] texels of the ramp texture


@dataclass
class TransformArea:  # ParticleTransform.cs:299-318
    Type: int = AreaType.None_
    Center: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Size: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    Falloff: float = 1.0
    Rotation: float = 0.0


class ParticleTransform:
    IsActive = True
    IsActive2 = True
    IsValid = True
    IsSpawner = False

    def pack(self, system: "ParticleSystem", now: float) -> Op:
        raise NotImplementedError


def _pack_area(area: Optional[TransformArea], strength: float, categoryFilter) -> Area:
    a = Area()
    if area is not None:
        a.AreaType = int(area.Type)
        a.AreaCenter[:] = [float(v) for v in area.Center]
        a.AreaSize[:] = [float(v) for v in area.Size]
        a.AreaFalloff = max(1.0, float(area.Falloff))
        a.AreaRotation = float(area.Rotation)
    else:
        a.AreaType = 0
        a.AreaFalloff = 1.0   # uniform left at its previous value by the reference; any finite value gives distance 0
        a.AreaSize[:] = [1.0, 1.0, 1.0]
    a.Strength = float(strength)
    cf = categoryFilter if categoryFilter is not None else (-9999.0, 9999.0)
    a.CategoryFilter[:] = [float(cf[0]), float(cf[1])]
    return a


def _time_divisor(cyclesPerSecond: Optional[float]) -> float:
    return float(F(VelocityConstantScale) / F(cyclesPerSecond)) if cyclesPerSecond is not None else -1.0


@dataclass
class FMA(ParticleTransform):  # Transforms.cs:16-49
    CyclesPerSecond: Optional[float] = 10
    PositionAdd: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    PositionMultiply: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    VelocityAdd: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    VelocityMultiply: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    Strength: float = 1.0
    CategoryFilter: Optional[Tuple[float, float]] = None
    Area: Optional[TransformArea] = None

    def pack(self, system, now):
        op = Op()
        op.kind = OP_FMA
        f = op.u.fma
        f.area = _pack_area(self.Area, self.Strength, self.CategoryFilter)
        f.TimeDivisor = _time_divisor(self.CyclesPerSecond)
        f.PositionAdd = Float4(*self.PositionAdd, 0)
        f.PositionMultiply = Float4(*self.PositionMultiply, 1)
        f.VelocityAdd = Float4(*self.VelocityAdd, 0)
        f.VelocityMultiply = Float4(*self.VelocityMultiply, 1)
        return op


@dataclass
class MatrixMultiply(ParticleTransform):  # Transforms.cs:51-80
    CyclesPerSecond: Optional[float] = 10
    Position: tuple = IDENTITY
    Velocity: tuple = IDENTITY
    Strength: float = 1.0
    CategoryFilter: Optional[Tuple[float, float]] = None
    Area: Optional[TransformArea] = None

    def pack(self, system, now):
        op = Op()
        op.kind = OP_MATRIX_MULTIPLY
        m = op.u.matrix
        m.area = _pack_area(self.Area, self.Strength, self.CategoryFilter)
        m.TimeDivisor = _time_divisor(self.CyclesPerSecond)
        m.PositionMatrix[:] = [float(v) for v in self.Position]
        m.VelocityMatrix[:] = [float(v) for v in self.Velocity]
        return op


@dataclass
class Noise(ParticleTransform):  # Transforms.cs:82-270
    IntervalUnit = 1000.0
    CyclesPerSecond: Optional[float] = 10
    PositionOffset: tuple = (-0.5, -0.5, -0.5, -0.5)
    PositionMinimum: tuple = (0.0, 0.0, 0.0, 0.0)
    PositionScale: tuple = (0.0, 0.0, 0.0, 0.0)
    VelocityOffset: tuple = (-0.5, -0.5, -0.5)
    VelocityMinimum: tuple = (0.0, 0.0, 0.0)
    VelocityScale: tuple = (1.0, 1.0, 1.0)
    SpeedOffset: float = -0.5
    SpeedMinimum: float = 0.0
    SpeedScale: float = 0.0
    Interval: float = 1000.0            # milliseconds between noise-field changes
    ReplaceOldVelocity: bool = True
    Strength: float = 1.0
    CategoryFilter: Optional[Tuple[float, float]] = None
    Area: Optional[TransformArea] = None
    Seed: int = 1

    def __post_init__(self):
        self._rng = np.random.default_rng(self.Seed)
        self.Reset()

    def _cycle(self):  # CycleUVs :149-154
        self.CurrentU, self.CurrentV = self.NextU, self.NextV
        self.NextU, self.NextV = float(self._rng.random()), float(self._rng.random())

    def Reset(self):  # :156-161
        self.LastUChangeWhen = 0.0
        self.NextU = self.NextV = 0.0
        self._cycle()

    def _auto_cycle(self, now: float, intervalSecs: float) -> float:  # AutoCycleUV :163-181
        if intervalSecs <= 0.01:
            return 0.0
        nextChangeWhen = self.LastUChangeWhen + intervalSecs
        if now >= nextChangeWhen:
            elapsed = now - nextChangeWhen
            self.LastUChangeWhen = now if elapsed >= intervalSecs else nextChangeWhen
            self._cycle()
        return float(F((now - self.LastUChangeWhen) / intervalSecs))

    def pack(self, system, now):
        op = Op()
        op.kind = OP_NOISE
        n = op.u.noise
        n.area = _pack_area(self.Area, self.Strength, self.CategoryFilter)
        n.TimeDivisor = _time_divisor(self.CyclesPerSecond)
        n.PositionOffset, n.PositionMinimum, n.PositionScale = Float4(*self.PositionOffset), Float4(*self.PositionMinimum), Float4(*self.PositionScale)
        n.VelocityOffset = Float4(*self.VelocityOffset, self.SpeedOffset)
        n.VelocityMinimum = Float4(*self.VelocityMinimum, self.SpeedMinimum)
        n.VelocityScale = Float4(*self.VelocityScale, self.SpeedScale)
        t = self._auto_cycle(float(F(now)), self.Interval / self.IntervalUnit)
        n.RandomnessOffset[:] = [float(F(self.CurrentU * 253)), float(F(self.CurrentV * 127))]
        n.NextRandomnessOffset[:] = [float(F(self.NextU * 253)), float(F(self.NextV * 127))]
        n.RandomnessTexel[:] = [float(F(1.0) / F(RandomnessTextureWidth)), float(F(1.0) / F(RandomnessTextureHeight))]
        n.FrequencyLerp = t
        n.ReplaceOldVelocity = 1.0 if self.ReplaceOldVelocity else 0.0
        return op


@dataclass
class Attractor:  # Transforms.cs:306-323
    Position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Radius: float = 1.0
    Strength: float = 1.0
    Type: int = AttractorType.Linear


@dataclass
class Gravity(ParticleTransform):  # Transforms.cs:303-372
    MaxAttractors = 16
    MaximumAcceleration: float = 8.0
    Attractors: List[Attractor] = field(default_factory=list)
    # The reference never sets Gravity.fx's CategoryFilter uniform, so the effect default (0, 0) applies: only
    # particles whose velocity.w (category / bounce delay) is exactly 0 are attracted.
    CategoryFilter: Tuple[float, float] = (0.0, 0.0)

    @property
    def IsValid(self):
        return len(self.Attractors) > 0

    def pack(self, system, now):
        if len(self.Attractors) > self.MaxAttractors:
            raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "Maximum number of attractors per instance is 16")
        op = Op()
        op.kind = OP_GRAVITY
        g = op.u.gravity
        g.AttractorCount = len(self.Attractors)
        g.MaximumAcceleration = float(self.MaximumAcceleration)
        g.CategoryFilter[:] = [float(self.CategoryFilter[0]), float(self.CategoryFilter[1])]
        for i, a in enumerate(self.Attractors):
            g.AttractorPositions[i] = Float4(*a.Position, 0)
            g.AttractorRadiusesAndStrengths[i] = Float4(a.Radius, a.Strength, int(a.Type), 0)
        return op


@dataclass
class Formula:  # Formula.cs: constant + (random + offset) * scale, or a circular variant
    Constant: tuple = (0.0, 0.0, 0.0)
    RandomScale: tuple = (0.0, 0.0, 0.0)
    Offset: tuple = (0.0, 0.0, 0.0)
    Type: int = FormulaType.Linear

    @property
    def Circular(self):
        return self.Type in (FormulaType.Spherical, FormulaType.Rectangular)


@dataclass
class Spawner(ParticleTransform):  # ParticleSpawner.cs:14-419
    IsSpawner = True
    MinRate: float = 0.0
    MaxRate: float = 0.0
    MaximumTotal: Optional[int] = None
    AlignVelocityAndPosition: bool = False
    AxisMask: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    Position: Formula = field(default_factory=Formula)
    Velocity: Formula = field(default_factory=lambda: Formula(Type=FormulaType.Spherical))
    Life: Tuple[float, float, float] = (1.0, 0.0, 0.0)      # Constant, RandomScale, Offset
    Category: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    ColorConstant: tuple = (1.0, 1.0, 1.0, 1.0)
    ColorRandomScale: tuple = (0.0, 0.0, 0.0, 0.0)
    ColorOffset: tuple = (0.0, 0.0, 0.0, 0.0)
    AlphaDiscardThreshold: float = 1.0
    PositionPostMatrix: tuple = IDENTITY
    VelocityPostMatrix: tuple = IDENTITY
    AdditionalPositions: List[Tuple[float, float, float]] = field(default_factory=list)
    PolygonRate: Optional[float] = None
    PolygonLoop: bool = True
    VelocityAlongPolygon: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    RatePerPosition: bool = True
    Seed: int = 1

    def __post_init__(self):
        self._rng = np.random.default_rng(self.Seed)
        self.RateError = 0.0
        self.TotalSpawned = 0
        self.Indices = (0, 0)

    @property
    def CountScale(self) -> int:  # :301-305
        return max(len(self.AdditionalPositions) + (1 if self.PolygonLoop else 0) if self.RatePerPosition else 1, 1)

    def BeginTick(self, now: float, deltaTimeSeconds: float) -> int:  # :152-189
        if not (self.IsActive and self.IsActive2):
            self.RateError = 0.0
            return 0
        minRate, maxRate = min(self.MinRate, self.MaxRate), self.MaxRate
        currentRate = ((float(self._rng.random()) * (maxRate - minRate)) + minRate) * self.CountScale * deltaTimeSeconds
        currentRate += self.RateError
        self.RateError = 0.0
        currentRate = self._adjust_current_rate(currentRate)
        if currentRate < 1:
            self.RateError = max(currentRate, 0.0)
            spawnCount = 0
        else:
            spawnCount = int(currentRate)
            self.RateError = currentRate - spawnCount
        if self.MaximumTotal is not None:
            remaining = self.MaximumTotal * self.CountScale - self.TotalSpawned
            if spawnCount > remaining:
                spawnCount = remaining
                self.RateError = 0.0
        return spawnCount

    def _adjust_current_rate(self, rate: float) -> float:  # AdjustCurrentRate :143-145
        return rate

    def EndTick(self, requested: int, actual: int):  # :191-194
        self.RateError += requested - actual
        self.TotalSpawned += actual

    def pack(self, system, now, chunk: int) -> Spawn:  # SetParameters :200-256, :376-403
        s = Spawn()
        s.chunk = chunk
        a, b = float(self._rng.random()), float(self._rng.random())
        s.RandomnessOffset[:] = [float(F(a * 253)), float(F(b * 127))]
        s.RandomnessTexel[:] = [float(F(1.0) / F(RandomnessTextureWidth)), float(F(1.0) / F(RandomnessTextureHeight))]
        count = 1 + len(self.AdditionalPositions)
        polygonRate = self.PolygonRate or 0.0
        if polygonRate >= 1:  # GetChunkSizeAndIndices :361-374
            c = count - 1 if (not self.PolygonLoop and count > 1) else count
            w = float(F(math.fmod(float(F(self.TotalSpawned / polygonRate)), float(c))))
        else:
            w = float(self.TotalSpawned % count)
        s.ChunkSizeAndIndices = Float4(system.Engine.Configuration.ChunkSize, self.Indices[0], self.Indices[1], w)
        lc, ls, lo = self.Life
        cc, cs, co = self.Category
        cfg = [Float4(*self.Position.RandomScale, ls), Float4(*self.Position.Offset, lo),
               Float4(*self.Velocity.Constant, cc), Float4(*self.Velocity.RandomScale, cs), Float4(*self.Velocity.Offset, co),
               Float4(*self.ColorConstant), Float4(*self.ColorRandomScale), Float4(*self.ColorOffset),
               Float4(*self.VelocityAlongPolygon, 0)]
        for i, c in enumerate(cfg):
            s.Configuration[i] = c
        s.FormulaTypes = Float4(int(self.Position.Type), int(self.Velocity.Type), 0, 0)
        s.AlignVelocityAndPosition = 1.0 if (self.AlignVelocityAndPosition and self.Position.Circular and self.Velocity.Circular) else 0.0
        s.AxisMask[:] = [float(v) for v in self.AxisMask]
        s.PositionMatrix[:] = [float(v) for v in self.PositionPostMatrix]
        s.VelocityMatrix[:] = [float(v) for v in self.VelocityPostMatrix]
        s.AttributeDiscardThreshold = float(F(self.AlphaDiscardThreshold) / F(255.0))
        self._source = None
        if count > MaxInlinePositions:  # SpawnParticlesFromPositionTexture (:291-295): PositionBuffer of (count + 127) / 128 * 128 texels (:306-352)
            buf = np.zeros(((count + 127) // 128 * 128, 4), dtype=np.float32)
            buf[0] = [*self.Position.Constant, lc]
            for i, ap in enumerate(self.AdditionalPositions):
                buf[i + 1] = [*ap, lc]
            src = SpawnSource()
            src.kind, src.position_count = _abi.SPAWN_POSITION_TEXTURE, buf.shape[0]
            src.positions = buf.ctypes.data
            self._position_buffer = buf
            system._spawn_keepalive.append(buf)   # the source holds a raw pointer: alive until the next plan_spawns
            self._source = src
        else:
            s.InlinePositionConstants[0] = Float4(*self.Position.Constant, lc)   # BeginTick :343-357
            for i, ap in enumerate(self.AdditionalPositions[:MaxInlinePositions - 1]):
                s.InlinePositionConstants[i + 1] = Float4(*ap, lc)
        s.PositionConstantCount = float(count)
        s.PolygonRate = float(polygonRate)
        s.PolygonLoop = 1.0 if self.PolygonLoop else 0.0
        return s


@dataclass
class FeedbackSpawner(Spawner):  # SpecialSpawners.cs:266-437 (derives from SpawnerBase: no additional positions / polygon)
    """Spawns from the particles of another system: position / velocity / colour constants come from a source particle."""
    SourceSystem: Optional["ParticleSystem"] = None
    SlidingWindowSize: Optional[int] = None
    SlidingWindowMargin: int = 0
    SpawnFromEntireWindow: bool = False
    InstanceMultiplier: int = 1
    AlignPositionConstant: bool = True
    SourceVelocityFactor: float = 0.0
    MultiplyLife: bool = False
    MultiplyColorConstant: bool = False
    SourceLifeRange: Tuple[float, float] = (0.0, 9999.0)

    IsFeedback = True

    def __post_init__(self):
        super().__post_init__()
        self.CurrentFeedbackSource = -1
        self.CurrentFeedbackSourceIndex = 0

    @property
    def CountScale(self) -> int:  # SpawnerBase :119-123
        return 1

    def BeginTickFeedback(self, system: "ParticleSystem", now: float, deltaTimeSeconds: float) -> Tuple[int, int]:
        """FeedbackSpawner.BeginTick (:325-403) -> (spawnCount, sourceChunk or -1)."""
        if self.InstanceMultiplier < 1:
            self.InstanceMultiplier = 1
        self.CurrentFeedbackSource = -1
        src = self.SourceSystem
        if src is None or src is system:      # :329-335
            return 0, -1
        spawnCount = Spawner.BeginTick(self, now, deltaTimeSeconds)
        if spawnCount < self.InstanceMultiplier and not self.SpawnFromEntireWindow:
            self.RateError += spawnCount       # AddError
            return 0, -1
        instances = spawnCount // self.InstanceMultiplier
        rounded = instances * self.InstanceMultiplier
        if rounded < spawnCount and rounded > 0:
            self.RateError += spawnCount - rounded
            spawnCount = rounded
        sourceChunk = src.PickSourceForFeedback(instances)
        if sourceChunk < 0:
            return 0, -1
        windowSize = self.SlidingWindowSize if self.SlidingWindowSize is not None else 999999
        available = src.AvailableForFeedback(sourceChunk)
        if src._chunk_no_longer_target[sourceChunk]:
            cw = src._spawn_target
            if cw >= 0:
                available += src.AvailableForFeedback(cw)
        windowedAvailable = min(available, windowSize)
        src.SkipFeedbackInput(sourceChunk, max(0, available - windowedAvailable))
        availableLessMargin = max(0, windowedAvailable - self.SlidingWindowMargin)
        spawnCount = min(spawnCount, availableLessMargin * self.InstanceMultiplier)
        spawnCount = min(spawnCount, src.AvailableForFeedback(sourceChunk) * self.InstanceMultiplier)
        self.CurrentFeedbackSource = sourceChunk
        self.CurrentFeedbackSourceIndex = src._chunk_consumed[sourceChunk]       # Chunk.FeedbackSourceIndex
        if self.SpawnFromEntireWindow:
            sourceCount = max(spawnCount // self.InstanceMultiplier, 1)
            maxOffset = availableLessMargin - sourceCount
            if maxOffset > 1:
                self.CurrentFeedbackSourceIndex += int(self._rng.integers(0, maxOffset))
        return spawnCount, sourceChunk

    def pack(self, system, now, chunk: int) -> Spawn:  # SetParameters :409-430 on top of SpawnerBase.SetParameters
        saved = self.AdditionalPositions, self.PolygonRate
        self.AdditionalPositions, self.PolygonRate = [], None
        try:
            s = Spawner.pack(self, system, now, chunk)
        finally:
            self.AdditionalPositions, self.PolygonRate = saved
        s.ChunkSizeAndIndices = Float4(system.Engine.Configuration.ChunkSize, self.Indices[0], self.Indices[1], 0)  # SpawnerBase :137-141
        s.Configuration[8] = Float4(0, 0, 0, 0)
        s.PolygonLoop = 0.0
        src = SpawnSource()
        src.kind = _abi.SPAWN_FEEDBACK
        src.source_system = self.SourceSystem.handle
        src.source_chunk = self.CurrentFeedbackSource
        src.FeedbackSourceIndex = float(self.CurrentFeedbackSourceIndex)
        src.InstanceMultiplier = float(self.InstanceMultiplier)
        src.SourceVelocityFactor = float(self.SourceVelocityFactor)
        src.AlignPositionConstant = 1.0 if self.AlignPositionConstant else 0.0
        src.MultiplyLife = 1.0 if self.MultiplyLife else 0.0
        src.MultiplyAttributeConstant = 1.0 if self.MultiplyColorConstant else 0.0
        src.SourceLifeRange[:] = [float(self.SourceLifeRange[0]), float(self.SourceLifeRange[1])]
        self._source = src
        return s


def _next_power_of_two(v: int) -> int:  # Squared.Util.Arithmetic.NextPowerOfTwo (un-vendored sq/Fracture): smallest 2^k >= v
    v = int(v)
    return 0 if v <= 0 else 1 << (v - 1).bit_length()


@dataclass
class PatternSpawner(Spawner):  # SpecialSpawners.cs:15-262 (derives from SpawnerBase)
    """One particle per `Divisor`-th pixel of `Texture` (uint8 [H, W, 4], SurfaceFormat.Color), coloured by the pixel."""
    Texture: Optional[np.ndarray] = None
    TextureTopLeftPx: Optional[Tuple[float, float]] = None
    TextureSizePx: Optional[Tuple[float, float]] = None
    MipBiasBase: float = -0.5
    Divisor: int = 1
    WholeSpawn: bool = False
    InstantInitialSpawn: bool = True
    MultiplyColorConstant: bool = True

    PartialSpawnAllowed = False      # :137-141

    def __post_init__(self):
        super().__post_init__()
        self.RowsSpawned = 0
        self.Divisor = min(max(int(self.Divisor), 1), 10)     # :45-52

    @property
    def IsValid(self) -> bool:
        return self.Texture is not None

    @property
    def DirectTextureSize(self) -> Tuple[float, float]:  # :76-97
        if self.Texture is None:
            return (0.0, 0.0)
        w, h = float(self.Texture.shape[1]), float(self.Texture.shape[0])
        if self.TextureSizePx is not None:
            if self.TextureSizePx[0] > 0:
                w = float(self.TextureSizePx[0])
            if self.TextureSizePx[1] > 0:
                h = float(self.TextureSizePx[1])
        if self.TextureTopLeftPx is not None:
            w, h = w - self.TextureTopLeftPx[0], h - self.TextureTopLeftPx[1]
        return (w, h)

    @property
    def ParticlesPerRow(self) -> int:   # :111-115
        return _next_power_of_two(int(self.DirectTextureSize[0]) // self.Divisor)

    @property
    def RowsPerInstance(self) -> int:   # :117-121
        return _next_power_of_two(int(self.DirectTextureSize[1]) // self.Divisor)

    @property
    def ParticlesPerInstance(self) -> int:
        return self.ParticlesPerRow * self.RowsPerInstance

    @property
    def CountScale(self) -> int:        # :129-136
        return self.ParticlesPerInstance if self.WholeSpawn else self.ParticlesPerRow

    def _adjust_current_rate(self, rate: float) -> float:   # AdjustCurrentRate :148-166
        if self.WholeSpawn and self.TotalSpawned == 0 and (self.MaximumTotal or 0) > 0 and rate >= 1 and self.InstantInitialSpawn:
            result = max(self.ParticlesPerInstance, rate)
            self.RateError += rate - result
            return result
        return rate

    def BeginTick(self, now: float, deltaTimeSeconds: float) -> int:  # :168-194
        if self.Texture is None:
            return 0
        spawnCount = Spawner.BeginTick(self, now, deltaTimeSeconds)
        minCount = self.ParticlesPerInstance if self.WholeSpawn else self.ParticlesPerRow
        if minCount <= 0:
            return 0
        requested = spawnCount
        if spawnCount < minCount:
            self.RateError += spawnCount
            return 0
        spawnCount = (spawnCount // minCount) * minCount
        self.RateError += requested - spawnCount
        return spawnCount

    def pack(self, system, now, chunk: int) -> Spawn:  # SetParameters :196-249 on top of SpawnerBase.SetParameters
        saved = self.AdditionalPositions, self.PolygonRate
        self.AdditionalPositions, self.PolygonRate = [], None
        try:
            s = Spawner.pack(self, system, now, chunk)
        finally:
            self.AdditionalPositions, self.PolygonRate = saved
        s.ChunkSizeAndIndices = Float4(system.Engine.Configuration.ChunkSize, self.Indices[0], self.Indices[1], 0)
        s.Configuration[8] = Float4(0, 0, 0, 0)
        s.PolygonLoop = 0.0
        tex = np.ascontiguousarray(self.Texture, dtype=np.uint8)
        th, tw = tex.shape[0], tex.shape[1]
        if self.WholeSpawn:
            currentRow, self.RowsSpawned = 0, 0
        else:
            currentRow = self.RowsSpawned % self.RowsPerInstance
            self.RowsSpawned += 1
        d = self.Divisor
        src = SpawnSource()
        src.kind = _abi.SPAWN_PATTERN
        src.pattern_texels, src.pattern_width, src.pattern_height = tex.ctypes.data, tw, th
        self._pattern_texels = tex
        system._spawn_keepalive.append(tex)   # the source holds a raw pointer: alive until the next plan_spawns
        src.StepWidthAndSizeScale = Float4(d, self.ParticlesPerRow, F(d) / F(tw), F(d) / F(th))
        baseX = baseY = F(0)
        if self.TextureTopLeftPx is not None:
            baseX, baseY = F(self.TextureTopLeftPx[0]) / F(tw), F(self.TextureTopLeftPx[1]) / F(th)
        # (currentRow * Divisor) / tex.Height is an INTEGER division in the reference (:231-234)
        src.YOffsetsAndCoordScale = Float4(currentRow, (currentRow * d) // th, d, d)
        src.TexelOffsetAndMipBias = Float4(F(-0.5) / F(tw) + baseX, F(-0.5) / F(th) + baseY, 0, F(math.log(d, 2)) + F(self.MipBiasBase))
        dts = self.DirectTextureSize
        src.CenteringOffset[:] = [float(F(dts[0]) * F(-0.5)), float(F(dts[1]) * F(-0.5))]
        src.MultiplyAttributeConstant = 1.0 if self.MultiplyColorConstant else 0.0
        self._source = src
        return s


class ParticleEngine:
    """ParticleEngine(content, coordinator, materials, configuration) -- ParticleEngine.cs:95-141."""

    def __init__(self, ctx: Optional[_abi.Context], configuration: Optional[ParticleEngineConfiguration] = None):
        self.ctx = ctx
        self.Configuration = configuration or ParticleEngineConfiguration()
        self.RandomnessTexture = generate_randomness_texture(self.Configuration.RandomSeed)


def generate_randomness_texture(seed: Optional[int]) -> np.ndarray:
    """GenerateRandomnessTexture (ParticleEngine.cs:495-544): 807x653 float4 of uniform [0,1) singles.  The reference
    seeds per-thread Xoshiro instances from the clock; here the table is a function of `seed` only."""
    rng = np.random.default_rng(0xB200 if seed is None else seed)
    return rng.random((RandomnessTextureHeight, RandomnessTextureWidth, 4), dtype=np.float32)


class ParticleSystem:
    """ParticleSystem(engine, configuration) -- ParticleSystem.cs:242-330 / Update :634-760."""
    MaxChunkCount = 64  # ParticleSystem.cs:49 (the capacity passed to the library may be larger, see `maxChunks`)
    LivenessCheckInterval = 4   # ParticleLiveness.cs:14

    def __init__(self, engine: ParticleEngine, configuration: Optional[ParticleSystemConfiguration] = None, maxChunks: Optional[int] = None):
        self.Engine = engine
        self.Configuration = configuration or ParticleSystemConfiguration()
        self.Transforms: List[ParticleTransform] = []
        self.ctx = engine.ctx
        self.ChunkSize = engine.Configuration.ChunkSize
        self.ChunkMaximumCount = self.ChunkSize * self.ChunkSize
        self.MaxChunks = maxChunks or self.MaxChunkCount
        self.CurrentFrameIndex = 0
        self.LastUpdateTimeSeconds: Optional[float] = None
        self.Now = 0.0
        self.TotalSpawnCount = 0
        self._chunk_next_offset: List[int] = []   # Chunk.NextSpawnOffset per live chunk (ParticleSystem.cs:148-240)
        self._chunk_total_spawned: List[int] = []      # Chunk.TotalSpawned
        self._chunk_consumed: List[int] = []           # Chunk.TotalConsumedForFeedback (== FeedbackSourceIndex)
        self._chunk_no_longer_target: List[bool] = []  # Chunk.NoLongerASpawnTarget
        self._chunk_is_feedback: List[bool] = []       # Chunk.IsFeedbackSource (set on the chunks a FeedbackSpawner fills)
        self._chunk_live_count: List[Optional[int]] = []   # LivenessInfo.Count (ParticleLiveness.cs:24-28); None until a count arrives
        self._chunk_dead_frames: List[int] = []            # LivenessInfo.DeadFrameCount
        self._chunk_reap: List[bool] = []                  # member of ChunksToReap
        self.DeadFrameThreshold = self.LivenessCheckInterval * 4   # ParticleLiveness.cs:22
        self._frames_until_liveness = 0                    # FramesUntilNextLivenessCheck
        self.IsClearPending = False
        self.ReapedChunkCount = 0
        self._spawn_target = -1
        self._feedback_spawn_target = -1               # CurrentFeedbackSpawnTarget
        self._feedback_source = -1                     # CurrentFeedbackSource
        self.last_sources = None                       # ilb_spawn_source list of the most recent plan_spawns (None: all inline)
        self._spawn_keepalive: list = []
        self.handle = None
        if self.ctx is not None:
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.ilb_particles_create(self.ctx.handle, self.ChunkSize, self.MaxChunks, C.byref(h)))
            self.handle = h
            rt = np.ascontiguousarray(engine.RandomnessTexture, dtype=np.float32)
            self.ctx.check(self.ctx.lib.ilb_particles_set_randomness(h, rt.ctypes.data_as(C.c_void_p), rt.shape[1], rt.shape[0]))
        self._life_ramp_uploaded = None

    def _sync_life_ramp(self) -> None:
        """MaybeSetLifeRampParameters (ParticleSystem.cs:911-925): binds Configuration.LifeRamp.Texture."""
        lr = self.Configuration.LifeRamp
        tex = lr.Texture if lr is not None else None
        if tex is self._life_ramp_uploaded or self.handle is None:
            return
        if tex is None:
            self.ctx.check(self.ctx.lib.ilb_particles_set_life_ramp(self.handle, None, 0, 0))
        else:
            arr = np.ascontiguousarray(tex, dtype=np.float32)
            self.ctx.check(self.ctx.lib.ilb_particles_set_life_ramp(self.handle, arr.ctypes.data_as(C.c_void_p), arr.shape[1], arr.shape[0]))
        self._life_ramp_uploaded = tex

    # ---- chunks -----------------------------------------------------------------------------------------------
    @property
    def LiveChunkCount(self) -> int:
        return len(self._chunk_next_offset)

    def _sync_chunk_lists(self) -> None:
        """Chunks that were registered by assigning `_chunk_next_offset` directly (tests that model a pre-filled user chunk)
        get the bookkeeping of a chunk whose NextSpawnOffset particles were all spawned."""
        for c in range(len(self._chunk_total_spawned), len(self._chunk_next_offset)):
            self._chunk_total_spawned.append(self._chunk_next_offset[c])
            self._chunk_consumed.append(0)
            self._chunk_no_longer_target.append(self._chunk_next_offset[c] >= self.ChunkMaximumCount)
            self._chunk_is_feedback.append(False)
        for c in range(len(self._chunk_live_count), len(self._chunk_next_offset)):
            self._chunk_live_count.append(None)   # GetLivenessInfo starts from TotalSpawned; nothing is known before the first count
            self._chunk_dead_frames.append(0)
            self._chunk_reap.append(False)

    def _create_chunk(self) -> int:  # CreateChunk (ParticleSystem.cs:520-545)
        if len(self._chunk_next_offset) >= self.MaxChunks:
            return -1
        self._sync_chunk_lists()
        self._chunk_next_offset.append(0)
        self._sync_chunk_lists()
        if self.handle:
            self.ctx.check(self.ctx.lib.ilb_particles_set_live_chunks(self.handle, len(self._chunk_next_offset)))
        return len(self._chunk_next_offset) - 1

    # ---- liveness and reaping (ParticleLiveness.cs:30-129, ParticleSystem.cs:529-545, :675, :702-714) -----------------------
    def Clear(self) -> None:
        """ParticleSystem.Clear (ParticleSystem.cs:529-545): every chunk is reaped at the top of the next Update."""
        self._sync_chunk_lists()
        self.IsClearPending = True
        for c in range(self.LiveChunkCount):
            self._chunk_reap[c] = True

    def _process_liveness(self, counts) -> None:
        """ProcessLatestLivenessInfo (ParticleLiveness.cs:46-78) for the chunks a liveness request covered."""
        self._sync_chunk_lists()
        for c, n in enumerate(counts):
            if c >= self.LiveChunkCount:
                break
            self._chunk_live_count[c] = int(n)
            self._chunk_dead_frames[c] = self._chunk_dead_frames[c] + 1 if n <= 0 else 0
            if self._chunk_dead_frames[c] >= self.DeadFrameThreshold:
                self._chunk_reap[c] = True

    def _poll_liveness(self, wait: bool = False) -> bool:
        if self.handle is None:
            return False
        counts = (C.c_int64 * max(self.MaxChunks, 1))()
        n = C.c_int(-1)
        self.ctx.check(self.ctx.lib.ilb_particles_poll_chunk_liveness(self.handle, counts, self.MaxChunks, C.byref(n), 1 if wait else 0))
        if n.value < 0:
            return False
        self._process_liveness([counts[i] for i in range(n.value)])
        return True

    def _reap_chunk(self, c: int) -> None:
        """Reap (ParticleLiveness.cs:121-129): the chunk leaves the ordered chunk list; later chunks keep their order."""
        if self.handle is not None:
            self.ctx.check(self.ctx.lib.ilb_particles_remove_chunk(self.handle, c))
        for lst in (self._chunk_next_offset, self._chunk_total_spawned, self._chunk_consumed, self._chunk_no_longer_target,
                    self._chunk_is_feedback, self._chunk_live_count, self._chunk_dead_frames, self._chunk_reap):
            del lst[c]
        fix = lambda i: -1 if i == c else (i - 1 if i > c else i)
        self._spawn_target, self._feedback_spawn_target, self._feedback_source = (fix(self._spawn_target), fix(self._feedback_spawn_target),
                                                                                  fix(self._feedback_source))
        self.ReapedChunkCount += 1

    def _update_live_count_and_reap(self) -> None:
        """UpdateLiveCountAndReapDeadChunks (ParticleLiveness.cs:80-106) + the pending Clear (ParticleSystem.cs:702-714)."""
        self._sync_chunk_lists()
        self._poll_liveness()
        for c in reversed(range(self.LiveChunkCount)):
            if self._chunk_reap[c]:
                self._reap_chunk(c)
        if self.IsClearPending:
            for c in reversed(range(self.LiveChunkCount)):
                self._reap_chunk(c)
            self.IsClearPending = False
            self.TotalSpawnCount = 0

    def Spawn(self, positions: np.ndarray, velocities: np.ndarray, attributes: np.ndarray) -> int:
        """Spawn(count, initializer) (ParticleSpawning.cs:61-113): fills NEW chunks with caller-provided state
        (arrays [n,4]); returns the first chunk index used."""
        n = positions.shape[0]
        per = self.ChunkMaximumCount
        first = None
        for start in range(0, n, per):
            c = self._create_chunk()
            if c < 0:
                raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "out of chunks")
            first = c if first is None else first
            m = min(per, n - start)
            bufs = []
            for src in (positions, velocities, attributes):
                b = np.zeros((per, 4), dtype=np.float32)
                b[:m] = src[start:start + m]
                bufs.append(b)
            self.ctx.check(self.ctx.lib.ilb_particles_upload_chunk(self.handle, c, *[b.ctypes.data_as(C.c_void_p) for b in bufs]))
            self._chunk_next_offset[c] = per   # user chunks are NoLongerASpawnTarget (ParticleSystem.cs:693-695)
            self._chunk_total_spawned[c] = m
            self._chunk_no_longer_target[c] = True
            self.TotalSpawnCount += per
        return first if first is not None else -1

    def WriteChunkBuffer(self, chunk: int, which: int, data: np.ndarray) -> None:
        """SetData on one texture of a chunk (0 PositionAndLife, 1 Velocity, 2 Attributes, 3 RenderColor, 4 RenderData)."""
        arr = np.ascontiguousarray(data, dtype=np.float32)
        if arr.shape != (self.ChunkMaximumCount, 4):
            raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, f"expected [{self.ChunkMaximumCount}, 4] texels, got {arr.shape}")
        self.ctx.check(self.ctx.lib.ilb_particles_upload_buffer(self.handle, chunk, which, arr.ctypes.data_as(C.c_void_p)))

    def ReadChunk(self, chunk: int):
        """Readback of one chunk (ParticleReadback.cs:59-61 / GetDataFast): (P, V, attributes, renderColor, renderData)."""
        per = self.ChunkMaximumCount
        outs = [np.empty((per, 4), dtype=np.float32) for _ in range(5)]
        self.ctx.check(self.ctx.lib.ilb_particles_download_chunk(self.handle, chunk, *[o.ctypes.data_as(C.c_void_p) for o in outs]))
        return tuple(outs)

    @property
    def LiveCount(self) -> int:
        out = C.c_int64(0)
        self.ctx.check(self.ctx.lib.ilb_particles_count_live(self.handle, C.byref(out)))
        return int(out.value)

    # ---- uniforms ---------------------------------------------------------------------------------------------
    def system_uniforms(self, deltaTimeSeconds: float) -> PsysUniforms:
        """Uniforms.ParticleSystem ctor (Uniforms.cs:208-235) + SetSystemUniforms (ParticleSystem.cs:547-575)."""
        cfg = self.Configuration
        u = PsysUniforms()
        cs = self.ChunkSize
        u.TexelAndSize = Float4(F(1) / F(cs), F(1) / F(cs), cfg.Size[0], cfg.Size[1])
        u.GlobalSettings = Float4(F(deltaTimeSeconds * VelocityConstantScale), cfg.Friction, cfg.MaximumVelocity, cfg.LifeDecayPerSecond)
        col = cfg.Collision
        if col is not None:
            u.CollisionSettings = Float4(col.EscapeVelocity, col.BounceVelocityMultiplier, col.Distance, col.LifePenalty)
        ar = cfg.AnimationRate
        u.AnimationRateAndRotationAndZToY = Float4(F(1.0) / F(ar[0]) if ar[0] != 0 else 0, F(1.0) / F(ar[1]) if ar[1] != 0 else 0,
                                                    1.0 if cfg.RotationFromVelocity else 0.0, cfg.ZToY)
        o = cfg.OpacityFromLife or 0
        if o != 0:
            b = Bezier4()
            b.A, b.B = Float4(1, 1, 1, 0), Float4(1, 1, 1, 1)
            b.RangeAndCount = Float4(0, F(1.0) / F(o), 2, 0)
            u.ColorFromLife = b
        else:
            u.ColorFromLife = clamped_bezier4(cfg.ColorFromLife)
        u.ColorFromVelocity = clamped_bezier4(cfg.ColorFromVelocity)
        u.SizeFromLife = clamped_bezier1(cfg.SizeFromLife)
        u.SizeFromVelocity = clamped_bezier1(cfg.SizeFromVelocity)
        lr = cfg.LifeRamp
        if lr is not None and lr.Texture is not None:   # MaybeSetLifeRampParameters ParticleSystem.cs:927-940
            rangeSize = max(F(lr.Maximum) - F(lr.Minimum), F(0.001))
            u.LifeRampSettings = Float4(F(lr.Strength) * (-1 if lr.Invert else 1), lr.Minimum, rangeSize, float(np.asarray(lr.Texture).shape[0]))
        else:
            u.LifeRampSettings = Float4(0, 0, 1, 1)
        self._sync_life_ramp()
        u.RotationFromLifeAndIndex[:] = [float(F(math.radians(cfg.RotationFromLife))), float(F(math.radians(cfg.RotationFromIndex)))]
        u.write_render_outputs = 1 if cfg.WriteRenderOutputs else 0
        if col is not None and col.DistanceField is not None:
            if col.DistanceFieldMaximumZ is None:  # ParticleSystem.cs:835-836
                raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "If a distance field is active, you must set DistanceFieldMaximumZ")
            u.has_collision_field = 1
            u.CollisionField = collision_field_uniforms(col.DistanceField, col.FullFieldAddressing)
        return u

    # ---- update -----------------------------------------------------------------------------------------------
    def _delta_time(self, now: float) -> float:  # ParticleSystem.cs:644-669 (variable-timestep branch)
        ups = self.Engine.Configuration.UpdatesPerSecond
        maxDelta = min(max(self.Engine.Configuration.MaximumUpdateDeltaTimeSeconds, 1 / 200.0), 10.0)
        tickUnit = 1.0 / min(max(ups if ups is not None else 60, 5), 200)
        dt = tickUnit
        if self.LastUpdateTimeSeconds is not None:
            dt = min(now - self.LastUpdateTimeSeconds, maxDelta)
        self.LastUpdateTimeSeconds = now
        return min(dt, maxDelta)

    # ---- feedback bookkeeping (ParticleSystem.Chunk :166-182, :234-238; ParticleSpawning.cs:246-264) ---------------
    def AvailableForFeedback(self, chunk: int) -> int:
        return self._chunk_total_spawned[chunk] - self._chunk_consumed[chunk]

    def SkipFeedbackInput(self, chunk: int, skipAmount: int) -> None:
        self._chunk_consumed[chunk] = min(self._chunk_consumed[chunk] + skipAmount, self._chunk_total_spawned[chunk])

    def PickSourceForFeedback(self, count: int) -> int:
        for c in range(self.LiveChunkCount):
            if self.AvailableForFeedback(c) >= count // 2 and not self._chunk_is_feedback[c]:
                self._feedback_source = c
                return c
        return -1

    def _pick_target_for_spawn(self, feedback: bool, count: int, partialSpawnAllowed: bool) -> int:  # PickTargetForSpawn :199-231
        chunk = self._feedback_spawn_target if feedback else self._spawn_target
        if chunk >= 0 and (self.ChunkMaximumCount - self._chunk_next_offset[chunk]) < (16 if partialSpawnAllowed else count):
            self._chunk_no_longer_target[chunk] = True
            chunk = -1
        if chunk < 0:
            chunk = self._create_chunk()
            if chunk < 0:
                return -1
            self._chunk_is_feedback[chunk] = feedback
        if feedback:
            self._feedback_spawn_target = chunk
        else:
            self._spawn_target = chunk
        return chunk

    def plan_spawns(self, now: float, dt: float):
        """RunSpawner for every active spawner (ParticleSpawning.cs:115-197); a partial spawn triggers the second
        RunSpawner pass (ParticleSystem.cs:733-740), which calls BeginTick again exactly like the reference.  The
        ilb_spawn_source of each spawn (None when every spawn is inline) is left in `self.last_sources`."""
        spawns, sources = [], []
        self._spawn_keepalive = []   # host arrays the ilb_spawn_source records of this plan point into (a spawner may pack twice per plan)
        self._sync_chunk_lists()
        for t in self.Transforms:
            if not t.IsSpawner or not (t.IsActive and t.IsActive2) or not t.IsValid:
                continue
            feedback = bool(getattr(t, "IsFeedback", False))
            for _pass in range(2):
                sourceChunk = -1
                if feedback:
                    requested, sourceChunk = t.BeginTickFeedback(self, now, dt)
                else:
                    requested = t.BeginTick(now, dt)
                if requested <= 0:
                    break
                spawnCount = min(requested, self.ChunkMaximumCount)
                partial = bool(getattr(t, "PartialSpawnAllowed", True))
                chunk = self._pick_target_for_spawn(feedback, spawnCount, partial)
                if chunk < 0:
                    break
                free = self.ChunkMaximumCount - self._chunk_next_offset[chunk]
                if spawnCount > free:          # ParticleSpawning.cs:144-149
                    if not partial:
                        break
                    spawnCount = free
                first = self._chunk_next_offset[chunk]
                t.Indices = (first, first + spawnCount - 1)
                self._chunk_next_offset[chunk] += spawnCount
                self.TotalSpawnCount += spawnCount
                if sourceChunk >= 0 and not t.SpawnFromEntireWindow:   # :163-170
                    t.SourceSystem._chunk_consumed[sourceChunk] += max(spawnCount // t.InstanceMultiplier, 1)
                spawns.append(t.pack(self, now, chunk))
                sources.append(getattr(t, "_source", None))
                t.EndTick(requested, spawnCount)
                self._chunk_total_spawned[chunk] += spawnCount
                if not (requested > spawnCount):
                    break
        self.last_sources = sources if any(x is not None for x in sources) else None
        return spawns

    def plan_ops(self, now: float):
        ops = [t.pack(self, now) for t in self.Transforms
               if not t.IsSpawner and t.IsActive and t.IsActive2 and t.IsValid]  # UpdateChunk :800-817
        return ops

    def Update(self, now: Optional[float] = None, deltaTimeSeconds: Optional[float] = None) -> None:
        """ParticleSystem.Update (ParticleSystem.cs:634): `now` plays TimeProvider.Seconds; pass deltaTimeSeconds to
        drive a fixed timestep."""
        self.CurrentFrameIndex += 1
        if now is None:
            now = self.Now + (deltaTimeSeconds if deltaTimeSeconds is not None else 1 / 60.0)
        dt = deltaTimeSeconds if deltaTimeSeconds is not None else self._delta_time(now)
        self.Now = now
        self.LastUpdateTimeSeconds = now
        self._update_live_count_and_reap()                 # ParticleSystem.cs:675
        computingLiveness = self._frames_until_liveness <= 0   # :716-720 (`FramesUntilNextLivenessCheck-- <= 0`)
        self._frames_until_liveness -= 1
        if computingLiveness:
            self._frames_until_liveness = self.LivenessCheckInterval
        spawns = self.plan_spawns(now, dt)
        ops = self.plan_ops(now)
        u = self.system_uniforms(dt)
        self.step_packed(u, spawns, ops, 1, self.last_sources)
        if computingLiveness and self.handle is not None and self.LiveChunkCount > 0:
            # ComputeLiveness (:752, ParticleLiveness.cs:131-141): counted behind this update, read back on a later frame
            self.ctx.check(self.ctx.lib.ilb_particles_request_chunk_liveness(self.handle))

    def step_packed(self, u: PsysUniforms, spawns, ops, steps: int = 1, sources=None) -> None:
        """`sources`: None (every spawn is inline) or one ilb_spawn_source / None per spawn (ilb_particles_step_sources)."""
        col = self.Configuration.Collision
        field_handle = col.DistanceField.handle if (col is not None and col.DistanceField is not None) else None
        self.ctx.check(self.ctx.lib.ilb_particles_set_collision_field(self.handle, field_handle))
        sp = (Spawn * max(len(spawns), 1))(*spawns)
        opa = (Op * max(len(ops), 1))(*ops)
        if sources is None:
            self.ctx.check(self.ctx.lib.ilb_particles_step(self.handle, C.byref(u), C.cast(sp, C.c_void_p), len(spawns),
                                                           C.cast(opa, C.c_void_p), len(ops), steps))
            return
        src = (SpawnSource * max(len(spawns), 1))(*[x if x is not None else SpawnSource() for x in sources])
        self.ctx.check(self.ctx.lib.ilb_particles_step_sources(self.handle, C.byref(u), C.cast(sp, C.c_void_p), C.cast(src, C.c_void_p),
                                                               len(spawns), C.cast(opa, C.c_void_p), len(ops), steps))

    # ---- rasterisation (N2) ---------------------------------------------------------------------------------------
    def render_params(self, width: int, height: int, blendState: str = "AlphaBlend", renderParams: Optional[ParticleRenderParameters] = None,
                      viewportPosition=(0.0, 0.0), viewportScale=(1.0, 1.0), clearColor=None, target_format: int = _abi.FORMAT_FLOAT4):
        """The uniforms of one ParticleSystem.Render draw: Uniforms.RasterizeParticleSystem (Uniforms.cs:238-290), the material
        choice and RenderingOptions of Render (ParticleSystem.cs:943-1039), RoundingPowerFromLife (:568-572)."""
        cfg, ap = self.Configuration, self.Configuration.Appearance
        rp = renderParams or ParticleRenderParameters()
        r = _abi.ParticleRender()
        r.width, r.height, r.target_format = int(width), int(height), target_format
        r.blend = {"AlphaBlend": _abi.BLEND_ALPHA, "Additive": _abi.BLEND_ADDITIVE, "Opaque": _abi.BLEND_OPAQUE}[blendState]
        tex = ap.Texture
        if tex is not None:
            th, tw = tex.shape[0], tex.shape[1]
            r.texture_filter = _abi.TEXTURE_LINEAR if ap.Bilinear else _abi.TEXTURE_POINT
            r.texture_width, r.texture_height = tw, th
            size = ap.SizePx if ap.SizePx is not None else (float(tw), float(th))
            ox, oy = F(ap.OffsetPx[0]) / F(tw), F(ap.OffsetPx[1]) / F(th)
            r.BitmapTextureRegion = Float4(ox, oy, ox + F(size[0]) / F(tw), oy + F(size[1]) / F(th))
            if ap.RelativeSize:
                r.SizeFactorAndPosition = Float4(F(size[0]) * F(0.5), F(size[1]) * F(0.5), rp.Origin[0], rp.Origin[1])
            else:
                r.SizeFactorAndPosition = Float4(1, 1, rp.Origin[0], rp.Origin[1])
        else:
            r.texture_filter = _abi.TEXTURE_NONE
            r.BitmapTextureRegion = Float4(0, 0, 1, 1)
            r.SizeFactorAndPosition = Float4(1, 1, rp.Origin[0], rp.Origin[1])
        r.Scale = Float4(rp.Scale[0], rp.Scale[1], 0, 0)
        g = cfg.GlobalColor
        r.GlobalColor = Float4(F(g[0]) * F(g[3]), F(g[1]) * F(g[3]), F(g[2]) * F(g[3]), g[3])
        r.ZFormula = Float4(*cfg.ZFormula)
        r.ZConfiguration = Float4(cfg.SizeFromZ, 0, 0, 0)
        r.RoundingPowerFromLife = clamped_bezier1(ap.RoundingPowerFromLife if ap.RoundingPowerFromLife is not None else BezierF(A=0.8, B=0.8, C=0.8, D=0.8))
        r.RenderingOptions = Float4(1 if ap.Rounded else 0, 1 if ap.DitheredOpacity else 0, 1 if ap.ColumnFromVelocity else 0,
                                    1 if ap.RowFromVelocity else 0)
        u = self.system_uniforms(0.0)
        r.TexelAndSize = u.TexelAndSize
        # the rasteriser's animation rate is the APPEARANCE's (ParticleSystem.cs:556-560 packs Appearance.AnimationRate)
        ar = ap.AnimationRate if (ap.AnimationRate[0] or ap.AnimationRate[1]) else cfg.AnimationRate
        r.AnimationRateAndRotationAndZToY = Float4(F(1.0) / F(ar[0]) if ar[0] != 0 else 0, F(1.0) / F(ar[1]) if ar[1] != 0 else 0,
                                                    u.AnimationRateAndRotationAndZToY.z, cfg.ZToY)
        r.ViewportPosition[:] = [float(viewportPosition[0]), float(viewportPosition[1])]
        r.ViewportScale[:] = [float(viewportScale[0]), float(viewportScale[1])]
        r.StippleFactor = float(rp.StippleFactor if rp.StippleFactor is not None else cfg.StippleFactor)
        if clearColor is not None:
            r.clear, r.ClearColor = 1, Float4(*clearColor)
        return r

    def Render(self, width: int, height: int, target: Optional[np.ndarray] = None, blendState: str = "AlphaBlend",
               renderParams: Optional[ParticleRenderParameters] = None, viewportPosition=(0.0, 0.0), viewportScale=(1.0, 1.0),
               clearColor=(0.0, 0.0, 0.0, 0.0)) -> np.ndarray:
        """ParticleSystem.Render (ParticleSystem.cs:943-1039) into a render target [H, W, 4] (float32, float16 or uint8).
        `target` = None renders over `clearColor`; an array is blended over and returned (a copy)."""
        fmt = _abi.FORMAT_FLOAT4
        if target is not None:
            target = np.array(target, copy=True, order="C")
            fmt = {np.dtype(np.float32): _abi.FORMAT_FLOAT4, np.dtype(np.float16): _abi.FORMAT_HALF4, np.dtype(np.uint8): _abi.FORMAT_RGBA8}[target.dtype]
            height, width = target.shape[0], target.shape[1]
        r = self.render_params(width, height, blendState, renderParams, viewportPosition, viewportScale,
                               clearColor if target is None else None, fmt)
        if target is None:
            target = np.empty((height, width, 4), dtype=np.float32)
        tex = self.Configuration.Appearance.Texture
        tex_ptr = None
        if tex is not None:
            tex = np.ascontiguousarray(tex, dtype=np.uint8)
            tex_ptr = tex.ctypes.data_as(C.c_void_p)
        self.ctx.check(self.ctx.lib.ilb_particles_render(self.handle, C.byref(r), tex_ptr, target.ctypes.data_as(C.c_void_p)))
        return target

    def RenderLayerDevice(self, d_layer: int, width: int, height: int, blendState: str = "AlphaBlend",
                          renderParams: Optional[ParticleRenderParameters] = None, viewportPosition=(0.0, 0.0), viewportScale=(1.0, 1.0)) -> None:
        """This system's chunks rendered over a TRANSPARENT float4 layer at the device pointer `d_layer` ([H, W, 4] float32): one
        rank's share of a multi-GPU ParticleSystem.Render (see composite_layers).  Asynchronous."""
        if self.Configuration.Appearance.Texture is not None:
            raise _abi.IlluminantError(_abi.ERR_UNSUPPORTED, "RenderLayerDevice renders untextured materials; pass the texture through ilb_particles_render_device")
        r = self.render_params(width, height, blendState, renderParams, viewportPosition, viewportScale, (0.0, 0.0, 0.0, 0.0), _abi.FORMAT_FLOAT4)
        self.ctx.check(self.ctx.lib.ilb_particles_render_device(self.handle, C.byref(r), None, C.c_void_p(int(d_layer))))

    def Dispose(self):
        if self.handle:
            self.ctx.lib.ilb_particles_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass


def composite_layers(ctx, layer_ptrs, width: int, height: int, rows, blendState: str, target_format: int, clearColor, target_ptrs) -> None:
    """Composites rows [rows[0], rows[1]) of the per-rank layers (device pointers in rank == draw order; peer-mapped pointers of
    other ranks are read over NVLink) onto `clearColor` (None: onto the contents of target_ptrs[0]) and stores the band into every
    target of `target_ptrs` (ilb_particles_composite_layers).  Asynchronous on the context's stream."""
    layers = (C.c_void_p * len(layer_ptrs))(*[C.c_void_p(int(p)) for p in layer_ptrs])
    targets = (C.c_void_p * len(target_ptrs))(*[C.c_void_p(int(p)) for p in target_ptrs])
    clear = Float4(*clearColor) if clearColor is not None else None
    blend = {"AlphaBlend": _abi.BLEND_ALPHA, "Additive": _abi.BLEND_ADDITIVE, "Opaque": _abi.BLEND_OPAQUE}[blendState]
    ctx.check(ctx.lib.ilb_particles_composite_layers(ctx.handle, layers, len(layer_ptrs), int(width), int(height), int(rows[0]), int(rows[1]), blend,
                                                     int(target_format), C.byref(clear) if clear is not None else None, targets, len(target_ptrs)))


def collision_field_uniforms(df: DistanceField, fullFieldAddressing: bool = False):
    """`new Uniforms.DistanceField(system.Configuration.Collision.DistanceField)` (ParticleTransform.cs:141-148): only the
    geometry members are initialised by that constructor, and NOTHING on the particle path sets DistanceFieldPacked1
    (its only writers are LightingRenderer.cs:1913 and :1933), so the update effect keeps the default (0,0,0,0):
    slicePosition = min(z, 0) * 0 -- particles collide against z-slice 0 of the field only.  That is the reference's
    behaviour and the default here; `fullFieldAddressing=True` (not in the reference) passes the field's real Packed1."""
    from .distance_field import RendererQualitySettings
    u = df.uniforms(RendererQualitySettings())
    u.ConeAndMisc = Float4(0, 0, 0, u.ConeAndMisc.w)             # Uniforms.cs:106
    u.StepAndMisc2 = Float4(0, 0, 1, u.StepAndMisc2.w)           # Uniforms.cs:107
    if not fullFieldAddressing:
        u.Packed1 = Float4(0, 0, 0, 0)
    return u
