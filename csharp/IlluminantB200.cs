// IlluminantB200.cs -- P/Invoke binding of libilluminant_b200.so for the reference's C# host (Squared.Illuminant).
//
// SOURCE ONLY: this image has no .NET toolchain, so the file is not compiled or tested here; it is the artefact a
// maintainer drops into Illuminant/ (see INTEGRATION.md for where the calls replace the draw submission).  Struct layouts
// mirror include/illuminant_b200.h field for field; tests/test_abi.py checks the equivalent ctypes mirror against the
// header's sizeof/offsetof, and the sizes asserted in the static constructor below are the same numbers.
// Convention follows the reference's own native binding (Squared.Nuklear/Squared.Nuklear/Nuklear.cs:10-12): Cdecl.
using System;
using System.Runtime.InteropServices;
using Microsoft.Xna.Framework;

namespace Squared.Illuminant.Native {
    public enum IlbStatus : int {
        OK = 0, InvalidArgument = -1, Cuda = -2, NoDevice = -3, InvalidOperation = -4, OutOfMemory = -5, Unsupported = -6
    }
    public enum IlbFormat : int { Float4 = 0, Half4 = 1, Rgba8 = 2 }
    public enum IlbOpKind : int { Gravity = 1, Noise = 2, FMA = 3, MatrixMultiply = 4 }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbDFUniforms {                       // ilb_df_uniforms: Uniforms.DistanceField (Uniforms.cs:79-108) + Packed1
        public Vector4 ConeAndMisc, TextureSliceAndTexelSize, StepAndMisc2, TextureSliceCount, Extent, Packed1;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbLightBatch {                       // ilb_light_batch: one LightTypeRenderState draw (LightingRenderer.cs:1149-1166)
        public int LightType, FirstVertex, VertexCount, RampTexture;   // RampTexture: 0 or an ilb_ramp_texture_create id
        public IlbDFUniforms DF;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbLightingFrame {                    // ilb_lighting_frame
        public int Width, Height, LightmapFormat, RowBegin, RowEnd, StencilCulling;
        public Vector4 EnvironmentZAndScale, EnvironmentZToY, GBufferTexelSizeAndMisc;
        public float GBufferViewportRelative, ViewportPositionX, ViewportPositionY, Reserved2;
        public Vector4 ClearColor;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbObstruction {                      // ilb_obstruction (LightObstruction.Vertex)
        public int Type;
        public Vector3 Center, Size;
        public Vector4 Rotation;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbPsysUniforms {                     // ilb_psys_uniforms
        public Vector4 GlobalSettings, CollisionSettings, TexelAndSize, AnimationRateAndRotationAndZToY;   // Uniforms.ParticleSystem
        public Uniforms.ClampedBezier4 ColorFromLife, ColorFromVelocity;                                   // Bezier.cs:589-600
        public Uniforms.ClampedBezier1 SizeFromLife, SizeFromVelocity;                                     // Bezier.cs:434-459
        public Vector4 LifeRampSettings;
        public Vector2 RotationFromLifeAndIndex;
        public int HasCollisionField, WriteRenderOutputs;
        public IlbDFUniforms CollisionField;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct IlbArea {                      // ilb_area
        public int AreaType;
        public Vector3 AreaCenter, AreaSize;
        public float AreaFalloff, AreaRotation, Strength;
        public Vector2 CategoryFilter;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct IlbGravity {                   // ilb_gravity
        public int AttractorCount;
        public float MaximumAcceleration;
        public Vector2 CategoryFilter;
        public fixed float AttractorPositions[16 * 4];
        public fixed float AttractorRadiusesAndStrengths[16 * 4];
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbNoise {                            // ilb_noise
        public IlbArea Area;
        public float TimeDivisor, FrequencyLerp, ReplaceOldVelocity, Reserved;
        public Vector2 RandomnessOffset, NextRandomnessOffset, RandomnessTexel;
        public Vector4 PositionOffset, PositionMinimum, PositionScale, VelocityOffset, VelocityMinimum, VelocityScale;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbFMA {                              // ilb_fma
        public IlbArea Area;
        public float TimeDivisor, Reserved0, Reserved1, Reserved2;
        public Vector4 PositionAdd, PositionMultiply, VelocityAdd, VelocityMultiply;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct IlbMatrixMultiply {                   // ilb_matrix_multiply
        public IlbArea Area;
        public float TimeDivisor, Reserved0, Reserved1, Reserved2;
        public Matrix PositionMatrix, VelocityMatrix;   // XNA Matrix is row-major M11..M44, row-vector convention
    }

    [StructLayout(LayoutKind.Explicit, Size = 16 + 528)]
    public struct IlbOp {                               // ilb_op: 16-byte header + union (largest member: ilb_gravity, 528 B)
        [FieldOffset(0)] public int Kind;
        [FieldOffset(16)] public IlbGravity Gravity;
        [FieldOffset(16)] public IlbNoise Noise;
        [FieldOffset(16)] public IlbFMA FMA;
        [FieldOffset(16)] public IlbMatrixMultiply Matrix;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct IlbSpawn {                     // ilb_spawn
        public int Chunk, Reserved0, Reserved1, Reserved2;
        public Vector4 ChunkSizeAndIndices;
        public fixed float Configuration[9 * 4];
        public Vector4 FormulaTypes;
        public fixed float InlinePositionConstants[4 * 4];
        public Matrix PositionMatrix, VelocityMatrix;
        public Vector2 RandomnessOffset, RandomnessTexel;
        public Vector3 AxisMask;
        public float AlignVelocityAndPosition, PositionConstantCount, PolygonRate, PolygonLoop, AttributeDiscardThreshold;
    }

    public static unsafe class B200 {
        public const string DllName = "illuminant_b200";
        const CallingConvention CC = CallingConvention.Cdecl;

        static B200 () {
            // same numbers tests/test_abi.py verifies for the ctypes mirror
            if (Marshal.SizeOf(typeof(LightVertex)) != 128 || Marshal.SizeOf(typeof(IlbDFUniforms)) != 96 ||
                Marshal.SizeOf(typeof(IlbLightBatch)) != 112 || Marshal.SizeOf(typeof(IlbOp)) != 544)
                throw new InvalidOperationException("illuminant_b200 struct layout mismatch");
        }

        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_abi_version ();
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_create (int deviceOrdinal, out IntPtr ctx);
        [DllImport(DllName, CallingConvention = CC)] public static extern void ilb_destroy (IntPtr ctx);
        [DllImport(DllName, CallingConvention = CC)] public static extern IntPtr ilb_last_error (IntPtr ctx);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_synchronize (IntPtr ctx);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_create (IntPtr ctx, int w, int h, void* rgba64, UIntPtr bytes, out IntPtr df);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_generate (IntPtr ctx, int w, int h, int sliceW, int sliceH, int sliceCount, ref IlbDFUniforms u, IlbObstruction* obstructions, int count, out IntPtr df);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_update_dynamic (IntPtr df, IntPtr staticDf, int sliceW, int sliceH, int sliceCount, ref IlbDFUniforms u, IlbObstruction* dynamicObstructions, int count);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_download (IntPtr df, void* rgba64, UIntPtr bytes);
        [DllImport(DllName, CallingConvention = CC)] public static extern void ilb_df_destroy (IntPtr df);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_gbuffer_upload (IntPtr ctx, int w, int h, int format, void* data);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_render_lighting (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, void* lightmapOut);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_render_lighting_frame (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, int gbufferWidth, int gbufferHeight, int gbufferFormat, void* gbuffer, void* lightmapOut);
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbParticleLightSource {   // ilb_particle_light_source
            public IntPtr System;
            public Vector4 LightProperties, MoreLightProperties, LightColor, LightSpecularColor;
            public IlbDFUniforms DF;
        }
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_render_lighting_frame_async (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, int gbufferWidth, int gbufferHeight, int gbufferFormat, void* gbuffer, void* lightmapOut, out ulong ticket);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_render_lighting_frame_wait (IntPtr ctx, ulong ticket);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_lighting_set_particle_lights (IntPtr ctx, IlbParticleLightSource* sources, int count);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_update_light_probes (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, Vector4* probePositions, Vector4* probeNormals, int probeCount, int outputFormat, void* probesOut);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_create (IntPtr ctx, int chunkSize, int maxChunks, out IntPtr psys);
        [DllImport(DllName, CallingConvention = CC)] public static extern void ilb_particles_destroy (IntPtr psys);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_set_randomness (IntPtr psys, Vector4* table, int w, int h);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_set_life_ramp (IntPtr psys, Vector4* texels, int w, int h);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_set_collision_field (IntPtr psys, IntPtr df);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_set_live_chunks (IntPtr psys, int count);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_upload_chunk (IntPtr psys, int chunk, Vector4* positionAndLife, Vector4* velocity, Vector4* attributes);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_download_chunk (IntPtr psys, int chunk, Vector4* positionAndLife, Vector4* velocity, Vector4* attributes, Vector4* renderColor, Vector4* renderData);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_step (IntPtr psys, ref IlbPsysUniforms uniforms, IlbSpawn* spawns, int spawnCount, IlbOp* ops, int opCount, int steps);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_count_live (IntPtr psys, out long count);

        // ---- "next" rows of SURVEY.md section 8f ------------------------------------------------------------------------
        // N3: lightmap resolve (Resolve.fx / HDR.fxh) and the luminance buffer behind RenderedLighting.TryComputeHistogram
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbResolve {                // ilb_resolve == what LightingResolveHandler._Before binds (LightingRenderer.cs:1464-1523)
            public int Width, Height, LightmapFormat, AlbedoFormat, OutputFormat, HdrMode;
            public float InverseScaleFactor, AlbedoIsSRGB, ResolveToSRGB, Offset, ExposureMinusOne, GammaMinusOne;
            public float MiddleGray, AverageLuminance, MaximumLuminanceSquared, WhitePoint;
            public float LightmapUVOffsetX, LightmapUVOffsetY, DitheringStrength, Reserved;
        }
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_resolve_lighting (IntPtr ctx, ref IlbResolve parameters, void* lightmapOrNull, void* albedoOrNull, void* output);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_resolve_lighting_device (IntPtr ctx, ref IlbResolve parameters, void* dLightmap, void* dAlbedoOrNull, void* dOutput);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_compute_luminance (IntPtr ctx, int width, int height, int lightmapFormat, void* lightmapOrNull, int level, float* outLuminance);

        // N4: spawner materials that read a source (SpawnParticlesFromPositionTexture, SpawnFeedbackParticles, SpawnPatternParticles)
        [StructLayout(LayoutKind.Sequential, Pack = 8)]
        public struct IlbSpawnSource {            // ilb_spawn_source, one per IlbSpawn of the same call
            public int Kind, PositionCount;       // 0 inline, 1 position texture, 2 feedback, 3 pattern
            public Vector4* Positions;            // Spawner.Temp4 (ParticleSpawner.cs:331-352)
            public IntPtr SourceSystem;           // FeedbackSpawner.SourceSystem.Instance's handle
            public int SourceChunk;
            public float FeedbackSourceIndex, InstanceMultiplier, SourceVelocityFactor;
            public float AlignPositionConstant, MultiplyLife, MultiplyAttributeConstant;
            public float SourceLifeRangeMin, SourceLifeRangeMax;
            public int Reserved;
            public byte* PatternTexels;           // PatternSpawner.Texture level 0 (SurfaceFormat.Color)
            public int PatternWidth, PatternHeight;
            public Vector4 StepWidthAndSizeScale, YOffsetsAndCoordScale, TexelOffsetAndMipBias;   // SpecialSpawners.cs:213-241
            public float CenteringOffsetX, CenteringOffsetY, Reserved2, Reserved3;
        }
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_step_sources (IntPtr psys, ref IlbPsysUniforms uniforms, IlbSpawn* spawns, IlbSpawnSource* sourcesOrNull, int spawnCount, IlbOp* ops, int opCount, int steps);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_upload_buffer (IntPtr psys, int chunk, int which, Vector4* data);

        // N2: ParticleSystem.Render (RasterizeParticleSystem.fx)
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbParticleRender {         // ilb_particle_render
            public int Width, Height, TargetFormat, Blend, TextureFilter, TextureWidth, TextureHeight, Clear;
            public Vector4 ClearColor;
            public Uniforms.RasterizeParticleSystem RasterizeSettings;   // 6 x Vector4, Uniforms.cs:238-290
            public Uniforms.ClampedBezier1 RoundingPowerFromLife;
            public Vector4 RenderingOptions, TexelAndSize, AnimationRateAndRotationAndZToY;
            public float ViewportPositionX, ViewportPositionY, ViewportScaleX, ViewportScaleY, StippleFactor, R0, R1, R2;
        }
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_render (IntPtr psys, ref IlbParticleRender parameters, void* textureOrNull, void* target);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_render_device (IntPtr psys, ref IlbParticleRender parameters, void* dTextureOrNull, void* dTarget);

        // device-pointer variants (interop with other CUDA code in the process, multi-GPU peer mappings) and introspection
        [DllImport(DllName, CallingConvention = CC)] public static extern IntPtr ilb_stream (IntPtr ctx);
        // page-locks caller memory (pinned arrays, MemoryMappedFile views) so that frame copies run by DMA; see INTEGRATION.md
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_host_register (IntPtr ctx, void* hostPtr, UIntPtr bytes);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_host_unregister (IntPtr ctx, void* hostPtr);
        [DllImport(DllName, CallingConvention = CC)] public static extern ulong ilb_launch_count (IntPtr ctx);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_create_device (IntPtr ctx, int w, int h, void* dRgba64, UIntPtr bytes, out IntPtr df);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_gbuffer_upload_device (IntPtr ctx, int w, int h, int format, void* dData);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_render_lighting_device (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, void* dLightmapOut);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_render_lighting_peers (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, void** dPeerLightmaps, int peerCount);
        [DllImport(DllName, CallingConvention = CC)] public static extern void* ilb_particles_device_buffer (IntPtr psys, int which);

        // ---- round 2 -----------------------------------------------------------------------------------------------------
        // scheduling knobs (ilb_option): never change results
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_set_option (IntPtr ctx, int option, int value);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_get_option (IntPtr ctx, int option, out int value);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_debug_detmath (IntPtr ctx, int function, float* x, float* result, int count);
        // a rank's row band of the G-buffer (multi-GPU); asynchronous UpdateLightProbes into a device buffer
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_gbuffer_upload_rows (IntPtr ctx, int w, int h, int format, int rowBegin, int rowEnd, void* rows);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_update_light_probes_device (IntPtr ctx, IntPtr df, ref IlbLightingFrame frame, IlbLightBatch* batches, int batchCount, LightVertex* vertices, int vertexCount, Vector4* probePositions, Vector4* probeNormals, int probeCount, int outputFormat, void* dProbesOut);
        // N1: height volumes and incremental slice updates (LightingRenderer.DistanceField.cs:80-260, :415-464)
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbHeightVolume {           // ilb_height_volume
            public int FirstEdge, EdgeCount;
            public float ZBase, Height;
            public float BoundsLeft, BoundsTop, BoundsRight, BoundsBottom;
        }
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_create_empty (IntPtr ctx, int w, int h, out IntPtr df);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_df_update_slices (IntPtr df, IntPtr staticDfOrNull, int sliceW, int sliceH, int sliceCount, ref IlbDFUniforms u, IlbObstruction* obstructions, int obstructionCount, IlbHeightVolume* volumes, int volumeCount, Vector4* edges, int edgeCount, int firstPhysicalSlice, int physicalSliceCount);
        // chunk liveness and reaping (ParticleLiveness.cs)
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_request_chunk_liveness (IntPtr psys);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_poll_chunk_liveness (IntPtr psys, long* counts, int capacity, out int count, int wait);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_remove_chunk (IntPtr psys, int chunk);
        // multi-GPU ParticleSystem.Render: per-rank transparent layers, composited band by band over peer mappings
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_particles_composite_layers (IntPtr ctx, void** dLayers, int layerCount, int width, int height, int rowBegin, int rowEnd, int blend, int targetFormat, Vector4* clearColorOrNull, void** dTargets, int targetCount);
        // N3: scaled / offset resolve (ResolveLighting drawn as a quad, LightingRenderer.cs:1537-1645)
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbResolvePlacement {       // ilb_resolve_placement
            public int TargetWidth, TargetHeight;
            public float PositionX, PositionY, ScaleX, ScaleY;
            public float AlbedoU0, AlbedoV0, AlbedoU1, AlbedoV1;
            public int AlbedoWidth, AlbedoHeight;
        }
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_resolve_lighting_placed (IntPtr ctx, ref IlbResolve parameters, ref IlbResolvePlacement placement, void* lightmapOrNull, void* albedoOrNull, void* target);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_resolve_lighting_placed_device (IntPtr ctx, ref IlbResolve parameters, ref IlbResolvePlacement placement, void* dLightmap, void* dAlbedoOrNull, void* dTarget);

        // N3: HDRConfiguration.Dithering and the LUT-blended resolve (LUTResolve.fx; IlluminantMaterials.SetLUTBlending)
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbDithering {              // ilb_dithering <- Squared.Render.DitheringSettings (FrameIndex = DeviceManager.FrameIndex)
            public float Strength, Unit, FrameIndex, BandSize, RangeMin, RangeMax;
        }
        [StructLayout(LayoutKind.Sequential, Pack = 4)]
        public struct IlbLutBlending {            // ilb_lut_blending <- LUTBlendingConfiguration + ColorLUT.Resolution / RowCount
            public int DarkResolution, BrightResolution, DarkRowCount, BrightRowCount;
            public float DarkLevel, NeutralBandSize, BrightLevel, PerChannel, LUTOnly;
            public Vector4 LUTOffsets;
            public float Reserved;
        }
        // LightSource.RampTexture -> SphereLightWithDistanceRamp (a 1x1 texture means "no ramp", LightingRenderer.cs:819-827: pass 0)
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_ramp_texture_create (IntPtr ctx, int width, int height, int format, void* texels, out int id);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_ramp_texture_destroy (IntPtr ctx, int id);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_set_dithering (IntPtr ctx, IlbDithering* settingsOrNull);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_resolve_lighting_lut (IntPtr ctx, ref IlbResolve parameters, ref IlbLutBlending lut, void* darkLut, void* brightLut, void* lightmapOrNull, void* albedo, void* output);
        [DllImport(DllName, CallingConvention = CC)] public static extern int ilb_resolve_lighting_lut_device (IntPtr ctx, ref IlbResolve parameters, ref IlbLutBlending lut, void* dDarkLut, void* dBrightLut, void* dLightmap, void* dAlbedo, void* dOutput);

        /// <summary>Maps an ilb_status to the exception type the reference throws at the same place.</summary>
        public static void Check (IntPtr ctx, int status) {
            if (status == 0)
                return;
            var message = Marshal.PtrToStringAnsi(ilb_last_error(ctx));
            switch ((IlbStatus)status) {
                case IlbStatus.InvalidArgument: throw new ArgumentException(message);
                case IlbStatus.InvalidOperation: throw new InvalidOperationException(message);   // ParticleSystem.cs:642, :836
                case IlbStatus.OutOfMemory: throw new OutOfMemoryException(message);
                case IlbStatus.Unsupported: throw new NotImplementedException(message);          // LightingRenderer.cs:203
                default: throw new Exception(message);
            }
        }
    }
}
