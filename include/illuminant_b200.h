/*
 * illuminant_b200.h -- C-ABI boundary of libilluminant_b200.so
 *
 * B200-native (sm_100a) replacement for the two data-parallel hot paths of
 * sq/Illuminant.  The reference has no FFI for these paths: the seam is its
 * *material layer* -- renderers fill blittable structs and enqueue one draw per
 * (effect file, technique).  Every entry point below replaces "enqueue draw(s)
 * with material X" and accepts the structs that cross at that draw bit-identically.
 * All `file:line` citations are relative to the reference tree (Illuminant/...).
 *
 * Conventions (mirrors the reference's own P/Invoke precedent,
 * Squared.Nuklear/Squared.Nuklear/Nuklear.cs:10-12 -- Cdecl, plain pointers):
 *   - every function returns 0 (ILB_OK) or a negative ilb_status; the message is
 *     available from ilb_last_error().  The C# shim maps codes to the exception
 *     types the reference throws (ParticleSystem.cs:642, :836).
 *   - the caller owns every host buffer; the library owns device memory behind
 *     opaque handles (mirrors Coordinator.DisposeResource, LightingRenderer.cs:658-684).
 *   - one context == one caller thread at a time (the reference issues all GPU work
 *     from the draw thread, LightingRenderer.cs:1030).  Work is asynchronous on the
 *     context's CUDA stream; entry points that fill HOST memory synchronise
 *     (like Texture2D.GetData), `_device` variants do not.
 *   - there is NO CPU fallback: without a CUDA device ilb_create fails.
 */
#ifndef ILLUMINANT_B200_H
#define ILLUMINANT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILB_API __attribute__((visibility("default")))
#define ILB_ABI_VERSION 1

typedef enum ilb_status {
    ILB_OK = 0,
    ILB_ERR_INVALID_ARGUMENT = -1,  /* ArgumentException / ArgumentOutOfRangeException */
    ILB_ERR_CUDA = -2,              /* device error (message carries cudaGetErrorString) */
    ILB_ERR_NO_DEVICE = -3,         /* no sm_100 device: the library never falls back to the CPU */
    ILB_ERR_INVALID_OPERATION = -4, /* InvalidOperationException (e.g. collision without a field, ParticleSystem.cs:836) */
    ILB_ERR_OUT_OF_MEMORY = -5,
    ILB_ERR_UNSUPPORTED = -6        /* feature outside the hot-path scope */
} ilb_status;

typedef struct ilb_float4 { float x, y, z, w; } ilb_float4;

typedef struct ilb_ctx ilb_ctx;   /* one CUDA device + stream                         */
typedef struct ilb_df ilb_df;     /* a DistanceField atlas resident in HBM            */
typedef struct ilb_psys ilb_psys; /* one ParticleSystem's chunk storage resident in HBM */

/* ------------------------------------------------------------------ context */

ILB_API int ilb_abi_version(void);
/* device_ordinal: CUDA device index (one process per GPU; rank r uses LOCAL_RANK). */
ILB_API int ilb_create(int device_ordinal, ilb_ctx** out_ctx);
ILB_API void ilb_destroy(ilb_ctx* ctx);
/* Last error message of this context; ctx == NULL returns the last creation error. */
ILB_API const char* ilb_last_error(const ilb_ctx* ctx);
/* Returns when everything queued on the context has run, frames in flight (ilb_render_lighting_frame_async) included. */
ILB_API int ilb_synchronize(ilb_ctx* ctx);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on. */
ILB_API void* ilb_stream(ilb_ctx* ctx);
/* Number of kernels this context has launched since creation (bench gpu_launches). */
ILB_API uint64_t ilb_launch_count(const ilb_ctx* ctx);
/* Page-locks `bytes` of caller memory at `host_ptr` (a GCHandle-pinned managed array, a native allocation, or a mapping of
 * shared memory -- MemoryMappedFile / shm_open) for every CUDA context of the process, so that the copies of
 * ilb_render_lighting_frame and ilb_gbuffer_upload_rows run asynchronously by DMA instead of being staged.  One node, one
 * process per GPU: when `lightmap_out` of every rank points into ONE shared mapping registered on every rank, each rank's band
 * of the frame goes from its GPU straight into the host frame of the consumer over the rank's own PCIe link -- the reassembled
 * frame reaches the host without passing through one GPU.  Unregister before the memory is freed or unmapped. */
ILB_API int ilb_host_register(ilb_ctx* ctx, void* host_ptr, size_t bytes);
ILB_API int ilb_host_unregister(ilb_ctx* ctx, void* host_ptr);

/* Scheduling knobs of the kernels, never results: light counts, discards, distance-field samples, march steps and particle
 * state are the same bits under every setting; ILB_OPT_LIGHT_CONST_BANK selects another instantiation of the light kernel, whose
 * smooth factors (ambient occlusion, specular) may differ in the last bit or two.  Defaults are the measured
 * best on B200; the environment variable ILB_OPT_<NAME> overrides a default at ilb_create. */
typedef enum ilb_option {
    ILB_OPT_LIGHT_CONCURRENT = 0,   /* 1: run the line-light pass and the sphere + directional pass side by side as co-resident
                                     * persistent grids instead of back to back (default 0: measured slower on B200, 8.4-9.8 ms
                                     * against 7.9 ms per C4 frame for every split of the SM, profiles/r2_light_sweep.jsonl) */
    ILB_OPT_LIGHT_LINE_CTAS = 1,    /* resident line-pass CTAs per SM while both passes run (default 2) */
    ILB_OPT_LIGHT_OTHER_CTAS = 2,   /* resident sphere + directional CTAs per SM while both passes run (default 2) */
    ILB_OPT_LIGHT_LINE_HELPERS = 3, /* extra line-pass CTAs per SM that start when the other pass has drained (default 1) */
    ILB_OPT_LIGHT_OTHER_HELPERS = 4,/* extra sphere + directional CTAs per SM for the opposite case (default 3) */
    ILB_OPT_LIGHT_PDL = 5,          /* 1: the sphere + directional pass is a programmatic dependent launch of the line-light pass:
                                     * its CTAs start in the idle slots of the first pass's last wave (default 1) */
    ILB_OPT_LIGHT_CONST_BANK = 6,   /* 1: frames of up to 256 lights keep their light records in the constant bank as well, so that the
                                     * per-pixel light loop re-reads a field where it uses it instead of holding the whole record in
                                     * registers across the cone trace (default 1) */
    ILB_OPT_LIGHT_SPLIT_BAND = 7,   /* 1: a device-to-device render of fewer than half of the frame's rows (one rank's band of a sharded
                                     * frame) is cut into two halves on two compute lanes, so that each half's last wave is filled by
                                     * the other half's CTAs (default 1) */
    ILB_OPT_LIGHT_TILE_ORDER = 8,   /* 1: the tiles of the sphere + directional pass are started heaviest first (by the number of sphere-light
                                     * quads that cover a tile, counted on the host from the frame's light list; equal tiles keep their
                                     * traversal order) so that a launch's last wave holds the cheap tiles (default 0: measured on B200, the
                                     * quad count predicts a tile's cost too poorly -- eight 270-row bands 7.87 ms against 7.86 ms, and the
                                     * whole C4 frame loses 2-D locality, 7.44 ms against 7.33 ms) */
    ILB_OPT_COUNT = 9
} ilb_option;
ILB_API int ilb_set_option(ilb_ctx* ctx, int option, int value);
ILB_API int ilb_get_option(const ilb_ctx* ctx, int option, int* out_value);

/* Diagnostic: evaluates the kernels' own device build of the deterministic sin / cos / acos (include/ilb_detmath.h) on
 * `count` host values, so that tests can compare it bit for bit with the host build of the same header.
 * function: 0 dm_sinf, 1 dm_cosf, 2 dm_acosf. */
ILB_API int ilb_debug_detmath(ilb_ctx* ctx, int function, const float* x, float* out, int count);

/* --------------------------------------------------------- distance field (L1) */

/* The `Uniforms.DistanceField` struct (Uniforms.cs:79-108, 5 x float4) followed by
 * `DistanceFieldPacked1` (LightingRenderer.cs:1933-1939).  Filled by the host exactly as
 * SetDistanceFieldParameters does (LightingRenderer.cs:1894-1940); quality-dependent
 * members make it per light batch / per particle system. */
typedef struct ilb_df_uniforms {
    ilb_float4 ConeAndMisc;              /* MaxConeRadius, DistanceFieldZOffset, OcclusionToOpacityPower, InvScaleFactorX */
    ilb_float4 TextureSliceAndTexelSize; /* 1/Columns, 1/Rows, 1/(VirtualWidth*Columns), 1/(VirtualHeight*Rows) */
    ilb_float4 StepAndMisc2;             /* StepLimit, MinimumLength, LongStepFactor, InvScaleFactorY */
    ilb_float4 TextureSliceCount;        /* Columns, Rows, maxValidZ, SliceCount */
    ilb_float4 Extent;                   /* VirtualWidth, VirtualHeight, VirtualDepth, MaximumEncodedDistance; x<=0 == no field */
    ilb_float4 Packed1;                  /* 1/Columns*(1/3 as float), SliceCount/Extent.z, maxValidZ, MinStepSize */
} ilb_df_uniforms;

/* Upload an `Rgba64` atlas in the raw layout of DistanceField.Save (SDF/DistanceField.cs:183-193):
 * texture_width*texture_height texels, row-major, 4 x uint16 UNORM per texel (r,g,b,a =
 * z-slices 3p..3p+3 of physical slice p, LightingRenderer.DistanceField.cs:356-361). */
ILB_API int ilb_df_create(ilb_ctx* ctx, int texture_width, int texture_height,
                          const uint16_t* rgba64, size_t bytes, ilb_df** out_df);
/* Same, from a device pointer on ctx's device (multi-GPU broadcast, on-GPU generation). */
ILB_API int ilb_df_create_device(ilb_ctx* ctx, int texture_width, int texture_height,
                                 const void* d_rgba64, size_t bytes, ilb_df** out_df);
ILB_API int ilb_df_download(ilb_df* df, uint16_t* rgba64, size_t bytes);
ILB_API void ilb_df_destroy(ilb_df* df);

/* "next" row N1 -- analytic obstruction rasterisation (LightObstruction.cs, DistanceFunction.fx:15-48,
 * LightingRenderer.DistanceField.cs:347-400): MAX-blend of encoded distances, 4 z per texel. */
typedef struct ilb_obstruction {
    int32_t type;          /* LightObstructionType: 1 Ellipsoid, 2 Box, 3 Cylinder, 4 Spheroid, 5 Octagon */
    float center[3];
    float size[3];
    float rotation[4];     /* quaternion (x,y,z,w) */
} ilb_obstruction;
ILB_API int ilb_df_generate(ilb_ctx* ctx, int texture_width, int texture_height,
                            int slice_width, int slice_height, int slice_count,
                            const ilb_df_uniforms* u, const ilb_obstruction* obstructions, int count,
                            ilb_df** out_df);
/* Height volumes (SDF/HeightVolume.cs, LightingRenderer.DistanceField.cs:185-260, Shaders/DistanceField.fx): an extruded
 * polygon [z_base, z_base + height].  `edges` is the reference's VertexDataTexture: one float4 (a.x, a.y, b.x, b.y) per polygon
 * edge j with a = p[j], b = p[wrap(j + 1)]; a volume owns edges [first_edge, first_edge + edge_count).  The volume's quad is
 * its polygon bounds expanded by DistanceLimit = 520 (LightingRenderer.cs:316) and is MAX-blended like the analytic
 * obstructions (LoadMaterials.cs:154-157).  The polygon distance itself (sdPolygonInit / sdPolygonVertex) lives in the
 * un-vendored sq/Fracture SDF2D.fxh and is restated from its published definition (Quilez, "2D distance functions", sdPolygon:
 * squared distance to the closest edge, sign flipped once per edge the horizontal ray from the point crosses). */
typedef struct ilb_height_volume {
    int32_t first_edge, edge_count;
    float z_base, height;
    float bounds[4];       /* polygon bounds: left, top, right, bottom */
} ilb_height_volume;

/* An atlas cleared to 0 (DistanceField.NeedClear, LightingRenderer.DistanceField.cs:51-55): every slice invalid. */
ILB_API int ilb_df_create_empty(ilb_ctx* ctx, int texture_width, int texture_height, ilb_df** out_df);

/* RenderDistanceFieldSliceTriplet (LightingRenderer.DistanceField.cs:80-152) for physical slices [first_physical_slice,
 * first_physical_slice + physical_slice_count): each is cleared (to 0, or to the static field's texels when static_df is given:
 * ClearDistanceField.fx:27-39), then the analytic obstructions and the height volumes are MAX-blended into it.  This is the
 * incremental update of RenderDistanceFieldPartition (:415-464): the host mirror calls it for at most
 * MaximumFieldUpdatesPerFrame slices per frame (LightingRenderer.Configuration.cs:88-91).  Texels of other slices are not
 * touched.  In place, asynchronous; derived data (expanded planes) is refreshed lazily. */
ILB_API int ilb_df_update_slices(ilb_df* df, const ilb_df* static_df, int slice_width, int slice_height, int slice_count,
                                 const ilb_df_uniforms* uniforms, const ilb_obstruction* obstructions, int obstruction_count,
                                 const ilb_height_volume* volumes, int volume_count, const ilb_float4* edges, int edge_count,
                                 int first_physical_slice, int physical_slice_count);

/* DynamicDistanceField (SDF/DistanceField.cs:248-310): the field a frame samples is the STATIC field with the dynamic
 * obstructions MAX-blended on top -- the slice is "cleared" to the static texture and only IsDynamic obstructions are
 * rasterised (LightingRenderer.DistanceField.cs:99-118, ClearDistanceField.fx:27-39).  Rewrites `df` in place from
 * `static_df` (same atlas size; no allocation, derived data is refreshed lazily), so it can run every frame. */
ILB_API int ilb_df_update_dynamic(ilb_df* df, const ilb_df* static_df, int slice_width, int slice_height, int slice_count,
                                  const ilb_df_uniforms* uniforms, const ilb_obstruction* dynamic_obstructions, int count);

/* ------------------------------------------------------------- G-buffer (L4) */

typedef enum ilb_format {
    ILB_FORMAT_FLOAT4 = 0, /* SurfaceFormat.Vector4     16 B */
    ILB_FORMAT_HALF4 = 1,  /* SurfaceFormat.HalfVector4  8 B */
    ILB_FORMAT_RGBA8 = 2   /* SurfaceFormat.Color        4 B (lightmap only) */
} ilb_format;

/* G-buffer texels in the encoding of GBufferShaderCommon.fxh:10-35 (GBuffer.cs:31-39).
 * data == NULL disables the G-buffer (flat ground, LightCommon.fxh:132-141). */
ILB_API int ilb_gbuffer_upload(ilb_ctx* ctx, int width, int height, int format, const void* data);
ILB_API int ilb_gbuffer_upload_device(ilb_ctx* ctx, int width, int height, int format, const void* d_data);
/* Rows [row_begin, row_end) of a width x height G-buffer, for a rank that shades only its row band: `rows` points at the
 * first texel of row_begin.  Asynchronous in stream order when `rows` is pinned host memory (which must then stay valid until
 * the next synchronising call); rows that were never uploaded hold zeros. */
ILB_API int ilb_gbuffer_upload_rows(ilb_ctx* ctx, int width, int height, int format, int row_begin, int row_end, const void* rows);

/* ----------------------------------------------------------- lighting (L2-L11) */

/* `LightVertex` (Vertices.cs:10-39): 8 x float4 = 128 B, Sequential, Pack=4 -- field order as declared
 * there.  Per-type packing: LightingRenderer.cs:1193-1219 (sphere), :1256-1307 (directional),
 * :1309-1337 (line); see SURVEY.md appendix A. */
typedef struct ilb_light_vertex {
    ilb_float4 LightPosition1, LightPosition2, LightPosition3;
    ilb_float4 LightProperties, MoreLightProperties, EvenMoreLightProperties;
    ilb_float4 Color1, Color2;
} ilb_light_vertex;

typedef enum ilb_light_type { /* LightSourceTypeID, LightSource.cs:12-21 */
    ILB_LIGHT_SPHERE = 1,
    ILB_LIGHT_DIRECTIONAL = 2,
    ILB_LIGHT_PARTICLE = 3,   /* never in a batch: see ilb_lighting_set_particle_lights */
    ILB_LIGHT_LINE = 4
} ilb_light_type;

/* One LightTypeRenderState (LightingRenderer.cs:801-837) == one instanced draw (:1149-1166):
 * a light type, the distance-field uniforms its material was given (quality is per batch) and a
 * range of LightVertex.  Batches are accumulated in array order (= the reference's draw order). */
typedef struct ilb_light_batch {
    int32_t light_type;   /* ilb_light_type */
    int32_t first_vertex;
    int32_t vertex_count;
    int32_t ramp_texture; /* 0, or an ilb_ramp_texture_create id: LightTypeRenderStateKey.RampTexture (LightingRenderer.cs:50, :150-158) */
    ilb_df_uniforms df;   /* Extent.x <= 0 == rendered without a distance field */
} ilb_light_batch;

/* LightSource.RampTexture (LightSource.cs:182-193): a sphere-light batch with a ramp texture is drawn with the
 * SphereLightWithDistanceRamp material (Shaders/SphereLight.fx:48-87, SphereLightCore.fxh:99-119, :160-199): the light's rgb is
 * RampTexture(preTraceOpacity, (atan2(dy, dx) + EvenMoreLightProperties.z) * EvenMoreLightProperties.w).rgb * coneOpacity
 * instead of the scalar opacity; RampTextureSampler is LINEAR, U CLAMP, V WRAP (RampCommon.fxh:5-12).  A 1x1 texture means
 * "no ramp" in the reference (LightingRenderer.cs:819-827): pass ramp_texture = 0 for it.  texels: width*height of `format`
 * (ILB_FORMAT_RGBA8 or ILB_FORMAT_FLOAT4), host memory.  Ramp textures on directional lights (DirectionalLightWithRamp) and on
 * light probes are outside the hot-path scope: ILB_ERR_UNSUPPORTED. */
ILB_API int ilb_ramp_texture_create(ilb_ctx* ctx, int width, int height, int format, const void* texels, int32_t* out_id);
ILB_API int ilb_ramp_texture_destroy(ilb_ctx* ctx, int32_t id);

/* Per-frame uniforms of the light pass. */
typedef struct ilb_lighting_frame {
    int32_t width, height;            /* lightmap size in pixels (render size) */
    int32_t lightmap_format;          /* ilb_format: HALF4 (HighQuality, LightingRenderer.cs:477-479), RGBA8, or FLOAT4 (parity tests) */
    int32_t row_begin, row_end;       /* rows [row_begin,row_end) are rendered: 0,height for a whole frame; a band per GPU */
    int32_t stencil_culling;          /* Configuration.StencilCulling: mask pixels like UpdateMaskFromGBuffer (GBufferMask.fx:26-44) */
    ilb_float4 EnvironmentZAndScale;  /* GroundZ, MaximumZ, RenderScale.x, RenderScale.y   (Uniforms.cs:14-77) */
    ilb_float4 EnvironmentZToY;       /* ZToYMultiplier, InvZToYMultiplier, LightOcclusion, 0 */
    ilb_float4 GBufferTexelSizeAndMisc; /* 1/w, 1/h (0,0 = G-buffer disabled), ViewportScale.x, .y (LightingRenderer.GBuffer.cs:520-534) */
    float GBufferViewportRelative;    /* 0 / 1 */
    float ViewportPosition[2];        /* view transform position incl. ScaleCompensation offset (LightingRenderer.cs:711-724) */
    float reserved2;
    ilb_float4 ClearColor;            /* Environment.Ambient * intensityScale, w zeroed in fullbright mode (:1013-1016) */
} ilb_lighting_frame;

/* RenderLighting (LightingRenderer.cs:917, replacing :1112-1168): clear to ClearColor, then additively
 * accumulate every light of every batch.  lightmap_out is HOST memory, width*(row_end-row_begin) texels of
 * lightmap_format, row row_begin first.  Synchronous. */
ILB_API int ilb_render_lighting(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                const ilb_light_batch* batches, int batch_count,
                                const ilb_light_vertex* vertices, int vertex_count,
                                void* lightmap_out);
/* One whole host-to-host frame in one call: the G-buffer a host-side rasteriser produced for this frame (the input
 * LightingRenderer.RenderLighting reads through GBufferTexelSizeAndMisc, LightingRenderer.GBuffer.cs:520-534) goes up,
 * the frame is shaded, the lightmap comes down -- software-pipelined over row bands on three CUDA streams, so the
 * copies hide behind the kernels.  Equivalent to ilb_gbuffer_upload + ilb_render_lighting (bit-identical lightmap).
 * The G-buffer must have the frame's size and be screen-aligned (GBufferViewportRelative == 0).  The overlap needs PINNED
 * (page-locked) gbuffer / lightmap_out buffers -- cudaHostAlloc / cudaHostRegister, in C# a GCHandle-pinned array registered
 * once at start-up; with pageable buffers the call is still correct (and all kernels are queued before the first download), but
 * every copy is staged through the driver and blocks the caller.  Synchronous: returns when lightmap_out is complete. */
ILB_API int ilb_render_lighting_frame(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                      const ilb_light_batch* batches, int batch_count,
                                      const ilb_light_vertex* vertices, int vertex_count,
                                      int gbuffer_width, int gbuffer_height, int gbuffer_format, const void* gbuffer,
                                      void* lightmap_out);
/* The same frame without the final wait: the call returns as soon as the frame is queued, *out_ticket names it, and
 * ilb_render_lighting_frame_wait(ctx, ticket) returns when its lightmap_out is complete (tickets complete in order; 0 = nothing
 * was queued, e.g. an empty row band).  A renderer that double-buffers its lit frame queues frame n + 1 before it waits for
 * frame n: consecutive frames of the same geometry are then chained band by band on the device -- the G-buffer rows of band b
 * of frame n + 1 go up as soon as the kernels of band b of frame n have run, its kernels start as soon as that band's lightmap
 * rows have gone down -- so the fill and drain of one frame's pipeline hide behind its neighbours.  Until the wait returns the
 * caller must leave `gbuffer` and `lightmap_out` alone (frames in flight need distinct lightmap_out buffers; both must be
 * page-locked for the call to be asynchronous at all).  Every other lighting, G-buffer or resolve entry point of the context
 * first waits for the frames in flight; light probes and particle systems do not. */
ILB_API int ilb_render_lighting_frame_async(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                            const ilb_light_batch* batches, int batch_count,
                                            const ilb_light_vertex* vertices, int vertex_count,
                                            int gbuffer_width, int gbuffer_height, int gbuffer_format, const void* gbuffer,
                                            void* lightmap_out, uint64_t* out_ticket);
ILB_API int ilb_render_lighting_frame_wait(ilb_ctx* ctx, uint64_t ticket);
/* Same with a DEVICE output pointer; asynchronous on ilb_stream(ctx). */
ILB_API int ilb_render_lighting_device(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                       const ilb_light_batch* batches, int batch_count,
                                       const ilb_light_vertex* vertices, int vertex_count,
                                       void* d_lightmap_out);
/* Peer-fused variant: each finished lightmap texel is stored to the same offset of every buffer in
 * d_peer_lightmaps[0..peer_count) (peer-mapped device pointers, e.g. over NVLink), so the all-gather of
 * row bands happens inside the kernel.  Buffers are FULL frames (width*height); only this band is written. */
ILB_API int ilb_render_lighting_peers(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                      const ilb_light_batch* batches, int batch_count,
                                      const ilb_light_vertex* vertices, int vertex_count,
                                      void* const* d_peer_lightmaps, int peer_count);

/* "next" row N4 -- ParticleLightSource (LightSource.cs:466-500): every live particle of `system` whose attribute colour
 * has alpha > 0 is a sphere light at the particle's position (ParticleLightVertexShader, ParticleLight.fx:16-82) with the
 * template's properties exactly as _ParticleLightBatchSetup sets them (LightingRenderer.cs:769-789):
 * LightProperties = (Radius, RampLength, RampMode, castsShadows && field ? 1 : 0), MoreLightProperties = (AO radius,
 * ShadowDistanceFalloff ?? -99999, FalloffYFactor, saturate(AO opacity)), LightColor = Template.Color,
 * LightSpecularColor = (SpecularColor, SpecularPower); `df` = the uniforms SetDistanceFieldParameters gives the batch.
 * The sources apply to every following ilb_render_lighting* call of the context (after the batches, in array order) until
 * replaced; count = 0 clears them.  The particle state is read on the device when the frame is rendered -- render before
 * ilb_particles_step to reproduce the reference's usePreviousData (LightingRenderer.cs:1137-1143).  StippleFactor is 1. */
typedef struct ilb_particle_light_source {
    ilb_psys* system;
    ilb_float4 LightProperties, MoreLightProperties, LightColor, LightSpecularColor;
    ilb_df_uniforms df;
} ilb_particle_light_source;
ILB_API int ilb_lighting_set_particle_lights(ilb_ctx* ctx, const ilb_particle_light_source* sources, int count);

/* UpdateLightProbes (LightingRenderer.LightProbes.cs:49-110): probe positions (xyz, opacity=1) and
 * normals (xyz, enableShadows) as uploaded by UpdateLightProbeTexture; result = probe_count texels
 * (HALF4 or FLOAT4), cleared to 0 then accumulated like the lightmap. HOST output, synchronous. */
ILB_API int ilb_update_light_probes(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                    const ilb_light_batch* batches, int batch_count,
                                    const ilb_light_vertex* vertices, int vertex_count,
                                    const ilb_float4* probe_positions, const ilb_float4* probe_normals,
                                    int probe_count, int output_format, void* probes_out);
/* Same, asynchronous: the texels go to a DEVICE buffer in stream order (the reference reads its probe target back later,
 * LightingRenderer.LightProbes.cs:88-110).  Positions / normals are host arrays, copied before the call returns. */
ILB_API int ilb_update_light_probes_device(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame,
                                           const ilb_light_batch* batches, int batch_count,
                                           const ilb_light_vertex* vertices, int vertex_count,
                                           const ilb_float4* probe_positions, const ilb_float4* probe_normals,
                                           int probe_count, int output_format, void* d_probes_out);

/* ------------------------------------------ "next" row N3: lightmap resolve / luminance */

typedef enum ilb_hdr_mode { /* HDRMode, LightingRenderer.HDR.cs:269-273 */
    ILB_HDR_NONE = 0,
    ILB_HDR_GAMMA_COMPRESS = 1,
    ILB_HDR_TONE_MAP = 2
} ilb_hdr_mode;

/* Everything LightingResolveHandler._Before binds for one resolve draw (LightingRenderer.cs:1464-1523) and the material
 * ResolveLighting picks (:1537-1591): hdr_mode + albedo select one of the six pixel shaders of Resolve.fx
 * ({,GammaCompressed,ToneMapped}LightingResolve{,WithAlbedo}PixelShader, Resolve.fx:66-217).  Values are the ones the
 * reference sets on the effect, clamps included (IlluminantMaterials.cs:81-137): the shim passes ExposureMinusOne =
 * clamp(Exposure, 1/256, 99999) - 1, GammaMinusOne = clamp(Gamma, 0.1, 4) - 1, WhitePoint (1 in mode NONE),
 * MaximumLuminanceSquared = clamp(MaximumLuminance)^2, InverseScaleFactor (0 means 1, :1469-1473).
 * Scope: the screen-aligned 1:1 resolve (RenderedLighting.Resolve with width/height = the lightmap's size, position 0;
 * with the LinearClamp sampler that fetches exactly one texel per pixel) -- LightmapUVOffset must be (0,0) and the
 * albedo has the lightmap's size.  ApplyDither (Resolve.fx:88) lives in the un-vendored sq/Fracture (DitherCommon.fxh);
 * the handler's default is Strength 0 (:1489-1494), i.e. the identity; a non-zero DitheringStrength (or ilb_set_dithering) dithers
 * under the convention stated at ilb_dithering.  LUT blending (LUTResolve.fx) is ilb_resolve_lighting_lut. */
typedef struct ilb_resolve {
    int32_t width, height;     /* lightmap (= output) size in pixels */
    int32_t lightmap_format;   /* ilb_format of the lightmap: HALF4, RGBA8 or FLOAT4 */
    int32_t albedo_format;     /* RGBA8 (SurfaceFormat.Color) or FLOAT4; ignored without albedo */
    int32_t output_format;     /* RGBA8 (backbuffer, UNORM round-to-nearest) or FLOAT4 (parity tests) */
    int32_t hdr_mode;          /* ilb_hdr_mode */
    float InverseScaleFactor;
    float AlbedoIsSRGB, ResolveToSRGB;          /* 0 / 1 */
    float Offset, ExposureMinusOne, GammaMinusOne;
    float MiddleGray, AverageLuminance, MaximumLuminanceSquared; /* GammaCompress (HDR.fxh:6-18) */
    float WhitePoint;                                            /* ToneMap (HDR.fxh:22-44) */
    float LightmapUVOffset[2];
    float DitheringStrength;
    float reserved;
} ilb_resolve;

/* One resolve draw, host to host: lightmap (NULL = the device-resident lightmap the context's most recent
 * ilb_render_lighting / ilb_render_lighting_frame call produced -- like RenderedLighting.Resolve, which never reads the
 * lightmap back) and albedo (NULL = the Resolve.fx shaders without albedo) are width*height texels; output is
 * width*height texels of output_format.  Synchronous. */
ILB_API int ilb_resolve_lighting(ilb_ctx* ctx, const ilb_resolve* params, const void* lightmap, const void* albedo, void* output);
/* Same with DEVICE pointers (16-byte aligned); asynchronous on ilb_stream(ctx). */
ILB_API int ilb_resolve_lighting_device(ilb_ctx* ctx, const ilb_resolve* params, const void* d_lightmap, const void* d_albedo,
                                        void* d_output);

/* Scaled / offset resolve: ResolveLighting draws the resolve as a bitmap quad at `Position` with `Scale` into a render target of
 * any size (LightingRenderer.cs:1537-1645; RenderedLighting.Resolve passes position, scale and the albedo region through).
 * With albedo the quad is the albedo region in texels times Scale; without, the lightmap's render size (params->width x
 * height) times Scale (the caller divides by RenderScale like :1635).  Every target pixel whose CENTRE lies inside the quad is
 * written (BlendState.Opaque), the others keep their contents.  Both textures are sampled LINEAR / CLAMP at the interpolated
 * texture coordinates (fp32 bilinear weights, the convention of the distance-field sampler), the lightmap at its coordinate plus
 * params->LightmapUVOffset, each clamped to its region (Resolve.fx:30-36, :47-53).  In the screen-aligned 1:1 case (Position 0,
 * Scale 1, target == lightmap size) the result equals ilb_resolve_lighting up to the fp32 rounding of the texture coordinates
 * (the neighbouring texel enters with a weight of a few 1e-7). */
typedef struct ilb_resolve_placement {
    int32_t target_width, target_height;
    float Position[2];
    float Scale[2];
    float AlbedoRegion[4];          /* u0, v0, u1, v1 (Bounds.Unit = 0, 0, 1, 1) */
    int32_t albedo_width, albedo_height;
} ilb_resolve_placement;
/* target: target_width x target_height texels of params->output_format, read (pixels outside the quad) and written. */
ILB_API int ilb_resolve_lighting_placed(ilb_ctx* ctx, const ilb_resolve* params, const ilb_resolve_placement* placement,
                                        const void* lightmap, const void* albedo, void* target);
ILB_API int ilb_resolve_lighting_placed_device(ilb_ctx* ctx, const ilb_resolve* params, const ilb_resolve_placement* placement,
                                               const void* d_lightmap, const void* d_albedo, void* d_target);
/* HDRConfiguration.Dithering (LightingRenderer.HDR.cs:212; the resolve handler binds it as the `Dithering` uniform with
 * FrameIndex = DeviceManager.FrameIndex, LightingRenderer.cs:1489-1497).  DitheringSettings and ApplyDither live in the
 * un-vendored sq/Fracture (Squared.Render, DitherCommon.fxh), so what this library computes is a stated CONVENTION, the same
 * in the kernels and in the oracle -- ordered dithering to multiples of 1 / Unit with the 17-periodic threshold pattern:
 *     t      = frac((2 * x + 7 * y + 23 * ((FrameIndex mod 4) + 0.5)) / 17)             (x, y = integer pixel coordinates, VPOS)
 *     rgb8   = rgb * Unit;  a = trunc(rgb8);  b = ceil(rgb8)
 *     q      = ((rgb8 - a) >= t * BandSize ? b : a) / Unit
 *     result = RangeMin <= rgb <= RangeMax ? lerp(rgb, q, Strength) : rgb                 (per channel)
 * Defaults for members left 0: Unit 255, BandSize 1, RangeMax 1.  Strength 0 is the identity (the handler's default). */
typedef struct ilb_dithering {
    float Strength, Unit, FrameIndex, BandSize, RangeMin, RangeMax;
} ilb_dithering;
/* Settings of every later resolve on this context; NULL restores the default (Strength 0).  A non-zero
 * ilb_resolve.DitheringStrength overrides Strength for that call. */
ILB_API int ilb_set_dithering(ilb_ctx* ctx, const ilb_dithering* settings);

/* LUT-blended resolve: {Screen,World}SpaceLUTBlendedLightingResolveWithAlbedo (Shaders/LUTResolve.fx:57-135), chosen by
 * ResolveLighting when a LUTBlendingConfiguration is given, an albedo is bound and the HDR mode is None
 * (LightingRenderer.cs:1558-1561, :1576-1579).  Members as IlluminantMaterials.SetLUTBlending packs them
 * (IlluminantMaterials.cs:139-149): LUTResolutionsAndRowCounts = (dark res, bright res, dark rows, bright rows), LUTLevels =
 * (DarkLevel, NeutralBandSize, BrightLevel); LUTOffsets is never set by the reference (0).  A ColorLUT texture is
 * SurfaceFormat.Color, (resolution * resolution) x (resolution * row_count) texels.  ReadLUT lives in the un-vendored
 * sq/Fracture (LUTCommon.fxh); the CONVENTION here is the usual strip layout -- slice b of row 0 occupies columns
 * [b * res, (b + 1) * res), red runs along x and green along y inside a slice at texel centres 0.5 + c * (res - 1), the two
 * slices around blue * (res - 1) are fetched LINEAR / CLAMP and blended by the fraction; LUTOffsets.xy / .zw are added to the
 * dark / bright (u, v). */
typedef struct ilb_lut_blending {
    int32_t dark_resolution, bright_resolution, dark_row_count, bright_row_count;
    float DarkLevel, NeutralBandSize, BrightLevel;
    float PerChannel, LUTOnly;                    /* 0 / 1 */
    float LUTOffsets[4];
    float reserved;
} ilb_lut_blending;
/* params->hdr_mode must be ILB_HDR_NONE and albedo must be given (the reference throws otherwise, :1593-1594).  dark_lut /
 * bright_lut: RGBA8 texels, HOST memory; the other pointers as for ilb_resolve_lighting.  Synchronous. */
ILB_API int ilb_resolve_lighting_lut(ilb_ctx* ctx, const ilb_resolve* params, const ilb_lut_blending* lut, const void* dark_lut,
                                     const void* bright_lut, const void* lightmap, const void* albedo, void* output);
/* Same with DEVICE pointers for all five buffers; asynchronous on ilb_stream(ctx). */
ILB_API int ilb_resolve_lighting_lut_device(ilb_ctx* ctx, const ilb_resolve* params, const ilb_lut_blending* lut, const void* d_dark_lut,
                                            const void* d_bright_lut, const void* d_lightmap, const void* d_albedo, void* d_output);

/* UpdateLuminanceBuffer + the mip chain TryComputeHistogram reads (LightingRenderer.cs:855-898, LightingRenderer.HDR.cs:154-186):
 * level 0 is (width/2) x (height/2) SurfaceFormat.Single texels, texel (x,y) = dot(lightmap texel (2x+1, 2y+1).rgb,
 * (0.299, 0.587, 0.144)) (CalculateLuminancePixelShader, Resolve.fx:219-234; point-sampled at the half-size target's pixel
 * centres; 0.144 is the reference's constant); level k+1 = 2x2 box filter of level k, size floor(size/2).  Writes level
 * `level` (level_width = width/2 >> level, level_height likewise; the reference's accuracyFactor, default 3) to HOST memory
 * out_luminance[level_width*level_height].  lightmap == NULL as for ilb_resolve_lighting.  Synchronous.  The histogram
 * itself (Histogram.cs) is host code in the reference and stays host code (illuminant_b200.hdr.Histogram). */
ILB_API int ilb_compute_luminance(ilb_ctx* ctx, int width, int height, int lightmap_format, const void* lightmap, int level,
                                  float* out_luminance);

/* ---------------------------------------------------------- particles (P1-P10) */

typedef struct ilb_bezier1 { ilb_float4 RangeAndCount, ABCD; } ilb_bezier1;         /* ClampedBezier1, Bezier.cs:434-459 */
typedef struct ilb_bezier4 { ilb_float4 RangeAndCount, A, B, C, D; } ilb_bezier4;   /* ClampedBezier4, Bezier.cs:589-600 */

/* Everything SetSystemUniforms / UpdateHandler._BeforeDraw bind for the update pass
 * (ParticleSystem.cs:547-575, ParticleTransform.cs:84-168, Uniforms.cs:197-236). */
typedef struct ilb_psys_uniforms {
    ilb_float4 GlobalSettings;     /* dt*1000, Friction, MaximumVelocity, LifeDecayPerSecond */
    ilb_float4 CollisionSettings;  /* EscapeVelocity, BounceVelocityMultiplier, Distance, LifePenalty */
    ilb_float4 TexelAndSize;       /* 1/ChunkSize, 1/ChunkSize, Size.x, Size.y */
    ilb_float4 AnimationRateAndRotationAndZToY;
    ilb_bezier4 ColorFromLife, ColorFromVelocity;
    ilb_bezier1 SizeFromLife, SizeFromVelocity;
    ilb_float4 LifeRampSettings;   /* strength (negative = inverted), minimum, range, index divisor (ParticleSystem.cs:911-941); x == 0: no ramp */
    float RotationFromLifeAndIndex[2]; /* radians (ParticleTransform.cs:155-158) */
    int32_t has_collision_field;   /* Configuration.Collision?.DistanceField != null -> UpdateWithDistanceField */
    int32_t write_render_outputs;  /* 1 = renderColor/renderData written (reference contract); 0 = 64 B/particle mode */
    ilb_df_uniforms CollisionField;/* Uniforms.DistanceField(collision field) (ParticleTransform.cs:141-148) */
} ilb_psys_uniforms;

typedef struct ilb_area { /* TransformArea as set by ParticleAreaTransform.SetParameters (ParticleTransform.cs:299-318) */
    int32_t AreaType;      /* 0 None, 1 Ellipsoid, 2 Box, 3 Cylinder, 4 Spheroid, 5 Octagon */
    float AreaCenter[3];
    float AreaSize[3];
    float AreaFalloff;     /* >= 1 */
    float AreaRotation;    /* scalar, broadcast to a float4 by the shader (FMA.fx:11,17) */
    float Strength;
    float CategoryFilter[2]; /* default (-9999, 9999) */
} ilb_area;

#define ILB_MAX_ATTRACTORS 16
typedef struct ilb_gravity { /* Gravity.fx:5-10, Transforms.cs:347-365 */
    int32_t AttractorCount;
    float MaximumAcceleration;
    float CategoryFilter[2];  /* the reference never sets it for Gravity: the effect default (0,0) applies */
    ilb_float4 AttractorPositions[ILB_MAX_ATTRACTORS];            /* xyz */
    ilb_float4 AttractorRadiusesAndStrengths[ILB_MAX_ATTRACTORS]; /* Radius, Strength, Type */
} ilb_gravity;

typedef struct ilb_noise { /* Noise.fx:5-19, Transforms.cs:243-268 */
    ilb_area area;
    float TimeDivisor;
    float FrequencyLerp;
    float ReplaceOldVelocity;
    float reserved;
    float RandomnessOffset[2], NextRandomnessOffset[2];
    float RandomnessTexel[2]; /* (1/807, 1/653) -- also used as the `rate` (Noise.fx:49-52) */
    ilb_float4 PositionOffset, PositionMinimum, PositionScale;
    ilb_float4 VelocityOffset, VelocityMinimum, VelocityScale;
} ilb_noise;

typedef struct ilb_fma { /* FMA.fx:4-13, Transforms.cs:38-45 */
    ilb_area area;
    float TimeDivisor;
    float reserved[3];
    ilb_float4 PositionAdd, PositionMultiply, VelocityAdd, VelocityMultiply;
} ilb_fma;

typedef struct ilb_matrix_multiply { /* MatrixMultiply.fx:4-12 ("next" row N4) */
    ilb_area area;
    float TimeDivisor;
    float reserved[3];
    float PositionMatrix[16], VelocityMatrix[16]; /* row-major XNA Matrix, row-vector convention mul(v, M) */
} ilb_matrix_multiply;

typedef enum ilb_op_kind { ILB_OP_GRAVITY = 1, ILB_OP_NOISE = 2, ILB_OP_FMA = 3, ILB_OP_MATRIX_MULTIPLY = 4 } ilb_op_kind;

typedef struct ilb_op { /* one active non-spawner ParticleTransform, in Transforms list order (ParticleSystem.cs:800-817) */
    int32_t kind;
    int32_t reserved[3];
    union {
        ilb_gravity gravity;
        ilb_noise noise;
        ilb_fma fma;
        ilb_matrix_multiply matrix;
    } u;
} ilb_op;

typedef struct ilb_spawn { /* one RunSpawner draw (ParticleSpawning.cs:115-197; uniforms ParticleSpawner.cs:200-256,361-403) */
    int32_t chunk;                  /* target chunk index */
    int32_t reserved[3];
    ilb_float4 ChunkSizeAndIndices; /* ChunkSize, first, last, polygon phase */
    ilb_float4 Configuration[9];
    ilb_float4 FormulaTypes;
    ilb_float4 InlinePositionConstants[4];
    float PositionMatrix[16], VelocityMatrix[16];
    float RandomnessOffset[2];
    float RandomnessTexel[2];
    float AxisMask[3];
    float AlignVelocityAndPosition;
    float PositionConstantCount;
    float PolygonRate;
    float PolygonLoop;
    float AttributeDiscardThreshold;
} ilb_spawn;

/* "next" row N4 -- the spawner materials that read a source besides the uniforms: SpawnParticlesFromPositionTexture (a Spawner
 * with more than MaxInlinePositions = 4 positions, ParticleSpawner.cs:268,331-352,376-384) and SpawnFeedbackParticles
 * (FeedbackSpawner, SpecialSpawners.cs:266-437).  One entry per ilb_spawn of the same call. */
typedef enum ilb_spawn_kind {
    ILB_SPAWN_INLINE = 0,            /* PS_Spawn, SpawnParticles.fx:10-30: positions from ilb_spawn.InlinePositionConstants */
    ILB_SPAWN_POSITION_TEXTURE = 1,  /* PS_SpawnFromPositionTexture, :32-52 */
    ILB_SPAWN_FEEDBACK = 2,          /* PS_SpawnFeedback, :54-120 */
    ILB_SPAWN_PATTERN = 3            /* PS_SpawnPattern, PatternSpawner.fx:21-96 (PatternSpawner, SpecialSpawners.cs:15-262) */
} ilb_spawn_kind;
typedef struct ilb_spawn_source {
    int32_t kind;                 /* ilb_spawn_kind */
    int32_t position_count;       /* POSITION_TEXTURE: texels of the PositionBuffer (ParticleSpawner.cs:306-319) */
    const ilb_float4* positions;  /* POSITION_TEXTURE: HOST array of (position, life) (:337-352); ilb_spawn.PositionConstantCount of them are addressed */
    ilb_psys* source_system;      /* FEEDBACK: SourceSystem.Instance -- another system of the same context (:333-335) */
    int32_t source_chunk;         /* FEEDBACK: the chunk PickSourceForFeedback chose (ParticleSpawning.cs:246-264); its PositionAndLife,
                                     Velocity and RenderColor are read as they are when the step runs */
    float FeedbackSourceIndex;    /* SpecialSpawners.cs:425 */
    float InstanceMultiplier;     /* >= 1 (:321-323) */
    float SourceVelocityFactor;
    float AlignPositionConstant, MultiplyLife, MultiplyAttributeConstant; /* 0 / 1 */
    float SourceLifeRange[2];
    int32_t reserved;
    /* PATTERN: one particle per `Divisor`-th pixel of a texture.  pattern_texels is the HOST level-0 image, SurfaceFormat.Color
     * (pattern_width*pattern_height*4 bytes, row-major); the mip chain the shader's tex2Dlod addresses through
     * TexelOffsetAndMipBias.w = log2(Divisor) + MipBiasBase is built on the device with a rounding 2x2 box filter.  Sampler as
     * declared in PatternSpawner.fx:11-19: LINEAR min/mag, POINT mip (level = floor(lod + 0.5) clamped), CLAMP addressing.
     * The four uniforms are the ones PatternSpawner.SetParameters computes (SpecialSpawners.cs:196-249);
     * MultiplyAttributeConstant (above) = MultiplyColorConstant. */
    const uint8_t* pattern_texels;
    int32_t pattern_width, pattern_height;
    ilb_float4 StepWidthAndSizeScale;  /* Divisor, ParticlesPerRow, Divisor / tex.Width, Divisor / tex.Height */
    ilb_float4 YOffsetsAndCoordScale;  /* currentRow, currentRow * Divisor / tex.Height, Divisor, Divisor */
    ilb_float4 TexelOffsetAndMipBias;  /* -0.5 / tex.Width + baseX, -0.5 / tex.Height + baseY, 0, log2(Divisor) + MipBiasBase */
    float CenteringOffset[2];
    float reserved2[2];
} ilb_spawn_source;

/* ParticleSystem storage: max_chunks chunks of chunk_size^2 particles (ParticleSystem.cs:73-240). */
ILB_API int ilb_particles_create(ilb_ctx* ctx, int chunk_size, int max_chunks, ilb_psys** out_psys);
ILB_API void ilb_particles_destroy(ilb_psys* psys);
/* The engine-wide randomness texture (ParticleEngine.cs:45-46, :495-544): width*height float4, row-major. */
ILB_API int ilb_particles_set_randomness(ilb_psys* psys, const ilb_float4* table, int width, int height);
/* LifeRampTexture (Configuration.Color.LifeRamp.Texture, ParticleSystem.cs:911-925; sampled POINT / U clamp / V wrap by
 * getRampedColorForLifeValueAndIndex, UpdateCommon.fxh:6-13,67-80): width*height float4 texels, row-major.  NULL removes
 * it (the reference then binds its white dummy ramp). */
ILB_API int ilb_particles_set_life_ramp(ilb_psys* psys, const ilb_float4* texels, int width, int height);
/* Configuration.Collision.DistanceField (may differ from the lighting field, SimpleParticles.cs:216-219). */
ILB_API int ilb_particles_set_collision_field(ilb_psys* psys, ilb_df* df);
/* Spawn(initializer)-style upload / readback of one chunk (ParticleWorkItems.cs:75-78, ParticleReadback.cs:59-61).
 * Arrays are chunk_size^2 float4, row-major; NULL pointers are skipped. */
ILB_API int ilb_particles_upload_chunk(ilb_psys* psys, int chunk, const ilb_float4* position_and_life,
                                       const ilb_float4* velocity, const ilb_float4* attributes);
ILB_API int ilb_particles_download_chunk(ilb_psys* psys, int chunk, ilb_float4* position_and_life,
                                         ilb_float4* velocity, ilb_float4* attributes,
                                         ilb_float4* render_color, ilb_float4* render_data);
/* SetData on one of a chunk's five textures (`which` as for ilb_particles_device_buffer: 0 PositionAndLife, 1 Velocity,
 * 2 Attributes, 3 RenderColor, 4 RenderData): chunk_size^2 float4 from HOST memory.  RenderColor / RenderData are outputs of the
 * update pass; uploading them restores a saved system for ilb_particles_render without stepping it. */
ILB_API int ilb_particles_upload_buffer(ilb_psys* psys, int chunk, int which, const ilb_float4* data);
/* Chunks [0, count) are live and updated by ilb_particles_step. */
ILB_API int ilb_particles_set_live_chunks(ilb_psys* psys, int count);
/* One ParticleSystem.Update (ParticleSystem.cs:634, replacing the RunSpawner/UpdateChunk loop :725-745):
 * spawners first, then for every live chunk the transforms in order, then UpdatePositions /
 * UpdateWithDistanceField.  `steps` repeats the same update (fixed uniforms) steps times; the spawn list is one tick's
 * spawns and is applied before the first of them only.  Asynchronous. */
ILB_API int ilb_particles_step(ilb_psys* psys, const ilb_psys_uniforms* uniforms,
                               const ilb_spawn* spawns, int spawn_count,
                               const ilb_op* ops, int op_count, int steps);
/* ilb_particles_step with a source per spawn: sources == NULL means every spawn is ILB_SPAWN_INLINE. */
ILB_API int ilb_particles_step_sources(ilb_psys* psys, const ilb_psys_uniforms* uniforms,
                                       const ilb_spawn* spawns, const ilb_spawn_source* sources, int spawn_count,
                                       const ilb_op* ops, int op_count, int steps);
/* Device pointers of the SoA state slabs (max_chunks*chunk_size^2 float4 each) for zero-copy consumers:
 * 0 PositionAndLife, 1 Velocity, 2 Attributes(Color), 3 RenderColor, 4 RenderData. */
ILB_API void* ilb_particles_device_buffer(ilb_psys* psys, int which);
/* Count particles with life > 0 in chunks [0, live) (CountLiveParticles.fx equivalent); synchronous. */
ILB_API int ilb_particles_count_live(ilb_psys* psys, int64_t* out_count);

/* Chunk liveness and reaping (ParticleLiveness.cs:14-129, ParticleSystem.cs:675, :702-714).  The reference counts the live
 * particles of every chunk with occlusion queries every LivenessCheckInterval frames, reads the counts back some frames later
 * and reaps chunks that stayed empty for DeadFrameThreshold checks, so that a continuous spawner never runs out of chunks.
 *   request: queues one count per live chunk behind the work already submitted; asynchronous.
 *   poll:    *out_count = number of chunks the last request covered and counts[0..*out_count) when its results have arrived
 *            (wait != 0 blocks until they have), -1 when there is none or it is still in flight.
 *   remove:  reaps `chunk`: the chunks behind it move down one slot, keeping their order (and their draw order); a request
 *            in flight is dropped, because it counted the old slots. */
ILB_API int ilb_particles_request_chunk_liveness(ilb_psys* psys);
ILB_API int ilb_particles_poll_chunk_liveness(ilb_psys* psys, int64_t* counts, int capacity, int* out_count, int wait);
ILB_API int ilb_particles_remove_chunk(ilb_psys* psys, int chunk);

/* ------------------------------------------ "next" row N2: particle rasterisation (ParticleSystem.Render) */

typedef enum ilb_blend {
    ILB_BLEND_ALPHA = 0,    /* BlendState.AlphaBlend (premultiplied): dst = src + dst * (1 - src.a) */
    ILB_BLEND_ADDITIVE = 1, /* BlendState.Additive: dst = src * src.a + dst */
    ILB_BLEND_OPAQUE = 2    /* BlendState.Opaque: dst = src */
} ilb_blend;
typedef enum ilb_texture_filter { ILB_TEXTURE_NONE = 0, ILB_TEXTURE_POINT = 1, ILB_TEXTURE_LINEAR = 2 } ilb_texture_filter;

/* One ParticleSystem.Render (ParticleSystem.cs:943-1039): every live chunk in order, every particle of the chunk in index order
 * (RenderChunk :876-907 draws quadCount instances), as the material RasterizeParticles{NoTexture,TexturePoint,TextureLinear} of
 * RasterizeParticleSystem.fx would: VS_PosVelAttr (:62-150) builds a rotated quad per live particle from PositionAndLife,
 * RenderData (size, rotation, |v|, category) and RenderColor; PS_NoTexture / PS_Texture / PS_TexturePoint (:190-254) shade it;
 * the output-merger blends in draw order.  Members are the uniforms of that draw:
 *   RasterizeSettings = `Uniforms.RasterizeParticleSystem` (Uniforms.cs:238-290): GlobalColor (premultiplied), BitmapTextureRegion,
 *   SizeFactorAndPosition, Scale, ZFormula (depth only: unused), ZConfiguration; RoundingPowerFromLife (ParticleSystem.cs:568-572);
 *   RenderingOptions = (Rounded, DitheredOpacity, ColumnFromVelocity, RowFromVelocity) (:1021-1028); TexelAndSize and
 *   AnimationRateAndRotationAndZToY from the system uniforms (Uniforms.cs:197-236).
 * Scope / conventions (the view transform, StippleReject and Dither64 live in the un-vendored sq/Fracture):
 *   - screen-space orthographic view: pixel = (world - ViewportPosition) * ViewportScale, pixel centres at +0.5; a pixel is
 *     covered when its centre lies in the quad's half-open unit square (-1 <= u < 1, -1 <= v < 1 in the quad's own frame);
 *   - StippleFactor must be 1 and DitheredOpacity 0 (ILB_ERR_UNSUPPORTED otherwise);
 *   - the sprite texture is SurfaceFormat.Color with one mip level (XNA Texture2D without mipmaps), CLAMP addressing;
 *   - blending accumulates in fp32 and converts to target_format once per pixel (the reference's output-merger rounds to the
 *     target's format after every quad). */
typedef struct ilb_particle_render {
    int32_t width, height;           /* render target size in pixels */
    int32_t target_format;           /* ilb_format: FLOAT4, HALF4 or RGBA8 */
    int32_t blend;                   /* ilb_blend */
    int32_t texture_filter;          /* ilb_texture_filter: NONE = RasterizeParticlesNoTexture */
    int32_t texture_width, texture_height;
    int32_t clear;                   /* 1: the target is cleared to ClearColor first; 0: blend over its contents */
    ilb_float4 ClearColor;
    ilb_float4 GlobalColor, BitmapTextureRegion, SizeFactorAndPosition, Scale, ZFormula, ZConfiguration;
    ilb_bezier1 RoundingPowerFromLife;
    ilb_float4 RenderingOptions;
    ilb_float4 TexelAndSize, AnimationRateAndRotationAndZToY;
    float ViewportPosition[2], ViewportScale[2];
    float StippleFactor;
    float reserved[3];
} ilb_particle_render;

/* texture: HOST Color texels (texture_width*texture_height*4 bytes) or NULL; target: HOST width*height texels of target_format,
 * read first when clear == 0, always written.  Synchronous. */
ILB_API int ilb_particles_render(ilb_psys* psys, const ilb_particle_render* params, const void* texture, void* target);
/* Same with DEVICE pointers (texture may be NULL); asynchronous on the context's stream apart from one 4-byte read-back of
 * the quad/tile pair count. */
ILB_API int ilb_particles_render_device(ilb_psys* psys, const ilb_particle_render* params, const void* d_texture, void* d_target);

/* Multi-GPU ParticleSystem.Render.  Chunk ranges are sharded over the ranks in draw order, so the frame the reference would draw
 * is every rank's chunks rendered over a TRANSPARENT float4 layer (ilb_particles_render_device with clear = 1, ClearColor = 0,
 * target_format = ILB_FORMAT_FLOAT4) and the layers composited in rank order: premultiplied "over" is associative (AlphaBlend),
 * additive blending is a sum (ILB_BLEND_OPAQUE layers carry no coverage: ILB_ERR_UNSUPPORTED).  This call composites rows
 * [row_begin, row_end) -- a rank's band -- of `layer_count` layers (device pointers, rank order; peer-mapped pointers of the
 * other ranks' layers are read over NVLink) onto clear_color (NULL: onto the current contents of d_targets[0]) and stores the
 * texels in target_format into every buffer of d_targets (peer-mapped full-frame targets: every rank ends with the whole
 * image).  Asynchronous; the caller orders it behind all ranks' layer renders (a barrier) like the lighting gather. */
ILB_API int ilb_particles_composite_layers(ilb_ctx* ctx, const void* const* d_layers, int layer_count, int width, int height,
                                           int row_begin, int row_end, int blend, int target_format,
                                           const ilb_float4* clear_color, void* const* d_targets, int target_count);

#ifdef __cplusplus
}
#endif
#endif /* ILLUMINANT_B200_H */
