// Host-side internals shared by the translation units of libilluminant_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/illuminant_b200.h"
#include "ilb_device.cuh"

struct ilb_df;
struct ilb_psys;

#define ILB_PIPELINE_BANDS 12

struct ilb_ctx {
    int device = -1;
    std::vector<ilb_df*> fields;      // children, destroyed with the context
    std::vector<ilb_psys*> systems;
    cudaStream_t stream = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    // G-buffer (L4)
    void* gbuffer = nullptr;
    bool gbuffer_owned = false;
    size_t gbuffer_capacity = 0;
    int gb_w = 0, gb_h = 0, gb_fmt = 0;
    // per-frame staging
    void* d_lights = nullptr;
    size_t d_lights_capacity = 0;
    void* h_lights[2] = {nullptr, nullptr};  // pinned, alternating (uploadLights)
    size_t h_lights_capacity[2] = {0, 0};
    cudaEvent_t ev_lights[2] = {nullptr, nullptr};
    int h_lights_next = 0;
    // what d_lights currently holds (uploadLights skips the copies of a frame whose flattened light list is unchanged)
    std::vector<char> lights_cache;
    const void* lights_cache_ptr = nullptr;
    unsigned long long lights_cache_const_epoch = 0;   // epoch of the constant bank that holds the same list; 0 = not in the bank
    void* d_lightmap = nullptr;  // staging for host-output entry points
    size_t d_lightmap_capacity = 0;
    void* d_probe_in = nullptr;
    size_t d_probe_in_capacity = 0;
    int lm_w = 0, lm_rows = 0, lm_fmt = -1;  // what d_lightmap holds after the last host-output frame (-1: nothing)
    // N3 resolve / luminance staging (resolve.cu)
    void* d_resolve_in = nullptr;   size_t d_resolve_in_capacity = 0;
    void* d_resolve_albedo = nullptr; size_t d_resolve_albedo_capacity = 0;
    void* d_resolve_out = nullptr;  size_t d_resolve_out_capacity = 0;
    void* d_resolve_lut = nullptr;  size_t d_resolve_lut_capacity = 0;   // dark + bright ColorLUT texels of a host-to-host LUT resolve
    // LightSource.RampTexture: id - 1 indexes `ramps`; `d_ramp_table` is the device copy of {texels, w, h} records (lighting.cu)
    struct RampTexture { float4* texels = nullptr; int w = 0, h = 0; };
    std::vector<RampTexture> ramps;
    void* d_ramp_table = nullptr; size_t d_ramp_table_capacity = 0; bool ramp_table_dirty = true;
    ilb_dithering dither = {0.0f, 255.0f, 0.0f, 1.0f, 0.0f, 1.0f};       // ilb_set_dithering; the handler's default
    void* d_luminance[2] = {nullptr, nullptr}; size_t d_luminance_capacity[2] = {0, 0};
    void* d_accum = nullptr;     // fp32 sums of the line-light pass (handed to / combined with the sphere + directional pass)
    size_t d_accum_capacity = 0;
    // concurrent lighting passes (lighting.cu, lightingLaunchRows): the other pass's fp32 sums, tile queues + arrival
    // counters, the streams of the second pass and of the two helper grids, fork / join events
    void* d_accum2 = nullptr;    size_t d_accum2_capacity = 0;
    void* d_tilework = nullptr;  size_t d_tilework_capacity = 0;
    cudaStream_t light_aux[3] = {};
    cudaEvent_t ev_light_fork = nullptr, ev_light_join[3] = {};
    int sm_count = 148;
    int opt[ILB_OPT_COUNT] = {};  // tuning knobs, ilb_set_option
    // heaviest-first tile orders of the sphere + directional pass (lighting.cu, tileOrderFor): a ring of cached orders keyed by the
    // frame geometry and the sphere lights' pixel rectangles, each with its own pinned staging and device copy; `used[lane]` is
    // recorded behind the last launch of that compute lane that reads the slot, and waited for before the slot is given away
    struct TileOrder {
        std::vector<int> key;
        unsigned* h = nullptr; unsigned* d = nullptr;
        size_t capacity = 0;
        cudaEvent_t used[2] = {nullptr, nullptr};
        unsigned long long stamp = 0;
    };
    TileOrder tile_orders[32];
    unsigned long long tile_order_clock = 0;
    // ParticleLightSources applied to every frame until replaced (ilb_lighting_set_particle_lights)
    std::vector<ilb_particle_light_source> particle_lights;
    void* d_plight_scratch = nullptr;
    size_t d_plight_scratch_capacity = 0;
    // host-to-host frame pipeline (ilb_render_lighting_frame): upload / download streams and per-band events
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // ilb_gbuffer_upload_rows copies on the upload stream (copy_in); it only waits for queued work that touches the same rows:
    // the lighting launches (and whole-buffer uploads) still in flight are remembered with the rows they read
    struct GBufferUser { cudaEvent_t done = nullptr; int row_begin = 0, row_end = 0; bool live = false; };
    GBufferUser gb_users[8];
    int gb_user_next = 0;
    cudaEvent_t ev_rows_uploaded[8] = {};
    int ev_rows_next = 0;
    cudaStream_t band_stream = nullptr;                        // second compute lane of the pipelined host-to-host frame
    cudaEvent_t ev_band_fork = nullptr, ev_band_join = nullptr;
    cudaEvent_t ev_in[ILB_PIPELINE_BANDS] = {}, ev_done[ILB_PIPELINE_BANDS] = {};
    // Frames in flight (ilb_render_lighting_frame_async): the next frame's bands are ordered behind the SAME band of the frame
    // before it -- its upload behind that band's kernels (ev_done[b]), its kernels behind that band's download (ev_down[b]) --
    // as long as the two frames share their geometry and nothing else touched the G-buffer in between (gb_generation is bumped by
    // every other entry point that writes or reads it).  ev_frame[ticket % 4] sits behind the last download of frame `ticket`.
    cudaEvent_t ev_down[ILB_PIPELINE_BANDS] = {}, ev_frame[4] = {};
    unsigned long long frame_ticket = 0, frame_waited = 0;     // frames enqueued / known complete
    unsigned long long gb_generation = 0, pipe_generation = ~0ull;
    int pipe_edges[ILB_PIPELINE_BANDS + 1] = {}, pipe_nb = 0, pipe_lanes = 0, pipe_w = 0, pipe_h = 0, pipe_gfmt = -1, pipe_lfmt = -1;
    void* pipe_gbuffer = nullptr; void* pipe_lightmap = nullptr;
};

#define ILB_MAX_VIRTUAL_SLICES 64

// one cached set of expanded planes (planes.cu), keyed by the addressing uniforms it was built for
struct ilb_df_planes {
    float key[10];
    uint64_t version = 0;   // ilb_df::version the planes were built from
    int nv = 0, sw = 0, sh = 0;
    float4* planes = nullptr;
    float4* vtab = nullptr;
    int pitch = 0;
};

struct ilb_df {
    ilb_ctx* ctx = nullptr;
    uint2* tex = nullptr;
    int tw = 0, th = 0;
    uint64_t version = 0;   // bumped when the atlas is rewritten in place (ilb_df_update_dynamic)
    std::vector<ilb_df_planes> planes;
};

#define ILB_RASTER_BUFFERS 11
struct ilb_psys {
    ilb_ctx* ctx = nullptr;
    int chunk_size = 0, max_chunks = 0, live_chunks = 0;
    size_t per_chunk = 0;
    float4* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // P, V, A, RC, RD
    float4* rng = nullptr;
    int rng_w = 0, rng_h = 0;
    ilb_df* field = nullptr;
    float4* life_ramp = nullptr;    // LifeRampTexture, float4 texels
    int life_ramp_w = 0, life_ramp_h = 0;
    float4* positions = nullptr;    // PositionBuffer of an ILB_SPAWN_POSITION_TEXTURE spawn
    size_t positions_capacity = 0;
    void* raster[ILB_RASTER_BUFFERS] = {};   // N2 rasteriser: counts, offsets, tile ranges, CUB temp, pairs x4, pair total, texture, target
    size_t raster_capacity[ILB_RASTER_BUFFERS] = {};
    uint8_t* pattern = nullptr;     // packed mip chain of an ILB_SPAWN_PATTERN spawn's texture
    size_t pattern_capacity = 0;
    float4* noise_table = nullptr;  // 2 * per_chunk float4, see noise_table_kernel
    float2* escape_table = nullptr; // per_chunk float2, see escape_table_kernel
    unsigned long long* d_count = nullptr;
    // per-chunk liveness, read back asynchronously (ilb_particles_request_chunk_liveness / _poll_chunk_liveness)
    unsigned long long* d_chunk_counts = nullptr;
    unsigned long long* h_chunk_counts = nullptr;  // pinned
    cudaEvent_t ev_chunk_counts = nullptr;
    int liveness_chunks = 0;
    bool liveness_pending = false;
    bool use_tma = false;  // ILB_PARTICLE_TMA=1 selects the TMA-staged persistent step kernel (measured 22 % slower: the
                           // chain is issue-bound and tile-lockstep adds barrier stalls; the direct kernel is the default)
    int sm_count = 148;
};

int ilb_fail(ilb_ctx* ctx, int code, const char* fmt, ...);
int ilb_cuda_fail(ilb_ctx* ctx, cudaError_t e, const char* what);
int ilb_reserve(ilb_ctx* ctx, void** ptr, size_t* capacity, size_t bytes, bool pinned_host);

#define ILB_CUDA(ctx, expr)                                             \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) return ilb_cuda_fail((ctx), _e, #expr);  \
    } while (0)

// Fills a DFGeometry from the reference uniform block; returns false when Extent.x <= 0 (no field).
bool ilb_make_df_geometry(const ilb_df* df, const ilb_df_uniforms& u, DFGeometry* out);

// planes.cu
int ilb_planes_attach(ilb_ctx* ctx, ilb_df* df, const ilb_df_uniforms& u, DFGeometry* g);
void ilb_planes_release(ilb_df* df);
// lighting.cu
int ilb_lighting_launch(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches,
                        int batch_count, const ilb_light_vertex* vertices, int vertex_count, void* const* d_outputs,
                        int output_count, bool outputs_are_full_frames);
int ilb_probes_launch(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches,
                      int batch_count, const ilb_light_vertex* vertices, int vertex_count, const ilb_float4* positions,
                      const ilb_float4* normals, int probe_count, int output_format, void* probes_out_host, void* d_probes_out);
int ilb_lighting_frame_from_host(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches,
                                 int batch_count, const ilb_light_vertex* vertices, int vertex_count, int gw, int gh, int gfmt,
                                 const void* gbuffer_host, void* lightmap_out_host, unsigned long long* out_ticket);
int ilb_lighting_frame_wait(ilb_ctx* ctx, unsigned long long ticket);
int ilb_frames_drain(ilb_ctx* ctx);   // waits for every frame in flight; a no-op when there is none
// resolve.cu (N3)
int ilb_resolve_launch(ilb_ctx* ctx, const ilb_resolve* params, const void* d_lightmap, const void* d_albedo, void* d_output);
int ilb_resolve_placed_launch(ilb_ctx* ctx, const ilb_resolve* params, const ilb_resolve_placement* placement, const void* d_lightmap,
                              const void* d_albedo, void* d_target);
int ilb_resolve_lut_launch(ilb_ctx* ctx, const ilb_resolve* params, const ilb_lut_blending* lut, const void* d_dark, const void* d_bright,
                           const void* d_lightmap, const void* d_albedo, void* d_output);
int ilb_luminance_launch(ilb_ctx* ctx, int width, int height, int lightmap_format, const void* d_lightmap, int level,
                         float* out_host);
size_t ilb_format_bytes(int format);
// api.cu: remembers that work queued on ctx->stream so far reads (or writes) G-buffer rows [row_begin, row_end)
int ilb_gbuffer_note_user(ilb_ctx* ctx, int row_begin, int row_end);
// raster.cu (N2)
int ilb_raster_launch(ilb_psys* psys, const ilb_particle_render* params, const void* d_texture, void* d_target);
void ilb_raster_release(ilb_psys* psys);
int ilb_raster_composite(ilb_ctx* ctx, const void* const* d_layers, int layer_count, int width, int height, int row_begin, int row_end,
                         int blend, int target_format, const ilb_float4* clear_color, void* const* d_targets, int target_count);
// dfgen.cu
int ilb_dfgen_launch(ilb_ctx* ctx, uint2* tex, const uint2* base, int tw, int th, int slice_w, int slice_h, int slice_count,
                     const ilb_df_uniforms* u, const ilb_obstruction* obs, int count, const ilb_height_volume* volumes, int volume_count,
                     const ilb_float4* edges, int edge_count, int first_physical, int physical_count);
// particles.cu
int ilb_particles_launch(ilb_psys* psys, const ilb_psys_uniforms* u, const ilb_spawn* spawns, const ilb_spawn_source* sources,
                         int spawn_count, const ilb_op* ops, int op_count, int steps);
int ilb_particles_count_launch(ilb_psys* psys, int64_t* out);
int ilb_particles_liveness_request(ilb_psys* psys);
int ilb_particles_liveness_poll(ilb_psys* psys, int64_t* counts, int capacity, int* out_count, int wait);
int ilb_particles_remove(ilb_psys* psys, int chunk);
