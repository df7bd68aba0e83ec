// C-ABI entry points of libilluminant_b200.so (include/illuminant_b200.h): argument validation, device-memory
// ownership and stream plumbing.  The kernels live in lighting.cu / particles.cu / dfgen.cu.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ilb_internal.h"

size_t ilb_format_bytes(int format);

#include <algorithm>
#include <unordered_set>

namespace {
std::mutex g_error_mutex;
std::string g_create_error;
// Live handles: a destroy call on a handle that is already gone (e.g. a field released after its context at
// interpreter shutdown) is a no-op instead of a use-after-free.
std::mutex g_live_mutex;
std::unordered_set<const void*> g_live;
void live_add(const void* p) { std::lock_guard<std::mutex> l(g_live_mutex); g_live.insert(p); }
bool live_take(const void* p) { std::lock_guard<std::mutex> l(g_live_mutex); return g_live.erase(p) != 0; }
bool live_has(const void* p) { std::lock_guard<std::mutex> l(g_live_mutex); return g_live.count(p) != 0; }
}  // namespace

int ilb_fail(ilb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    else {
        std::lock_guard<std::mutex> lock(g_error_mutex);
        g_create_error = buf;
    }
    return code;
}

int ilb_cuda_fail(ilb_ctx* ctx, cudaError_t e, const char* what) {
    const int code = (e == cudaErrorMemoryAllocation) ? ILB_ERR_OUT_OF_MEMORY : ILB_ERR_CUDA;
    return ilb_fail(ctx, code, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int ilb_reserve(ilb_ctx* ctx, void** ptr, size_t* capacity, size_t bytes, bool pinned_host) {
    if (*capacity >= bytes && *ptr) return ILB_OK;
    if (*ptr) {
        ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pinned_host) cudaFreeHost(*ptr); else cudaFree(*ptr);
        *ptr = nullptr;
        *capacity = 0;
    }
    const size_t want = bytes + bytes / 2;
    if (pinned_host) ILB_CUDA(ctx, cudaMallocHost(ptr, want));
    else ILB_CUDA(ctx, cudaMalloc(ptr, want));
    *capacity = want;
    return ILB_OK;
}

bool ilb_make_df_geometry(const ilb_df* df, const ilb_df_uniforms& u, DFGeometry* g) {
    memset(g, 0, sizeof(*g));
    if (!df || !(u.Extent.x > 0.0f)) return false;
    g->tex = df->tex;
    g->tw = df->tw; g->th = df->th;
    g->twf = (float)df->tw; g->thf = (float)df->th;
    g->inv_tw = 1.0f / (float)df->tw;
    g->zOffset = u.ConeAndMisc.y;
    g->ex = u.Extent.x; g->ey = u.Extent.y; g->ez = u.Extent.z; g->maxEnc = u.Extent.w;
    g->maxValidZ = u.Packed1.z; g->zToSlice = u.Packed1.y; g->invSliceCountXTimesOneThird = u.Packed1.x;
    g->sliceSizeX = u.TextureSliceAndTexelSize.x; g->sliceSizeY = u.TextureSliceAndTexelSize.y;
    g->texelSizeX = u.TextureSliceAndTexelSize.z; g->texelSizeY = u.TextureSliceAndTexelSize.w;
    g->invScaleX = u.ConeAndMisc.w; g->invScaleY = u.StepAndMisc2.w;
    g->sliceCount = u.TextureSliceCount.w;
    return true;
}

namespace {
// the device build of include/ilb_detmath.h exactly as the kernels include it (ilb_device.cuh)
__global__ void detmath_kernel(int function, const float* __restrict__ x, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    out[i] = function == 0 ? dm_sinf(v) : (function == 1 ? dm_cosf(v) : dm_acosf(v));
}
}  // namespace

int ilb_gbuffer_note_user(ilb_ctx* ctx, int row_begin, int row_end) {
    ctx->gb_generation++;   // a reader / writer of the G-buffer outside the frame pipeline (see ilb_ctx::ev_down)
    ilb_ctx::GBufferUser& u = ctx->gb_users[ctx->gb_user_next];
    ctx->gb_user_next = (ctx->gb_user_next + 1) % 8;
    if (!u.done) ILB_CUDA(ctx, cudaEventCreateWithFlags(&u.done, cudaEventDisableTiming));
    // a slot that is reused keeps the rows of the launch it replaces: this launch is later in the stream, so its event covers both
    if (u.live) { row_begin = std::min(row_begin, u.row_begin); row_end = std::max(row_end, u.row_end); }
    ILB_CUDA(ctx, cudaEventRecord(u.done, ctx->stream));
    u.row_begin = row_begin; u.row_end = row_end; u.live = true;
    return ILB_OK;
}

extern "C" {

int ilb_abi_version(void) { return ILB_ABI_VERSION; }

int ilb_debug_detmath(ilb_ctx* ctx, int function, const float* x, float* out, int count) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (function < 0 || function > 2 || count < 0 || (count > 0 && (!x || !out))) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad argument");
    if (count == 0) return ILB_OK;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    float* d = nullptr;
    ILB_CUDA(ctx, cudaMalloc(&d, sizeof(float) * 2 * (size_t)count));
    cudaError_t e = cudaMemcpyAsync(d, x, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        detmath_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(function, d, d + count, count);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + count, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return ilb_cuda_fail(ctx, e, "ilb_debug_detmath");
    return ILB_OK;
}

int ilb_create(int device_ordinal, ilb_ctx** out_ctx) {
    if (!out_ctx) return ilb_fail(nullptr, ILB_ERR_INVALID_ARGUMENT, "out_ctx is null");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return ilb_fail(nullptr, ILB_ERR_NO_DEVICE, "no CUDA device (%s); illuminant_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= count)
        return ilb_fail(nullptr, ILB_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device_ordinal, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_ordinal);
    if (e != cudaSuccess) return ilb_cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return ilb_fail(nullptr, ILB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device_ordinal, prop.major, prop.minor);
    e = cudaSetDevice(device_ordinal);
    if (e != cudaSuccess) return ilb_cuda_fail(nullptr, e, "cudaSetDevice");
    ilb_ctx* ctx = new ilb_ctx();
    ctx->device = device_ordinal;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return ilb_cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
    static const struct { const char* env; int value; } defaults[ILB_OPT_COUNT] = {
        {"ILB_OPT_LIGHT_CONCURRENT", 0}, {"ILB_OPT_LIGHT_LINE_CTAS", 2}, {"ILB_OPT_LIGHT_OTHER_CTAS", 2},
        {"ILB_OPT_LIGHT_LINE_HELPERS", 1}, {"ILB_OPT_LIGHT_OTHER_HELPERS", 3}, {"ILB_OPT_LIGHT_PDL", 1}, {"ILB_OPT_LIGHT_CONST_BANK", 1},
        {"ILB_OPT_LIGHT_SPLIT_BAND", 1}, {"ILB_OPT_LIGHT_TILE_ORDER", 0}};
    for (int i = 0; i < ILB_OPT_COUNT; i++) {
        const char* e = getenv(defaults[i].env);
        ctx->opt[i] = e ? atoi(e) : defaults[i].value;
    }
    live_add(ctx);
    *out_ctx = ctx;
    return ILB_OK;
}

int ilb_set_option(ilb_ctx* ctx, int option, int value) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (option < 0 || option >= ILB_OPT_COUNT) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "unknown option %d", option);
    if (value < 0 || value > 32) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "option %d: value %d outside [0,32]", option, value);
    ctx->opt[option] = value;
    return ILB_OK;
}

int ilb_get_option(const ilb_ctx* ctx, int option, int* out_value) {
    if (!ctx || !out_value || option < 0 || option >= ILB_OPT_COUNT) return ILB_ERR_INVALID_ARGUMENT;
    *out_value = ctx->opt[option];
    return ILB_OK;
}

void ilb_destroy(ilb_ctx* ctx) {
    if (!ctx || !live_take(ctx)) return;
    cudaSetDevice(ctx->device);
    ilb_frames_drain(ctx);
    cudaStreamSynchronize(ctx->stream);
    while (!ctx->fields.empty()) ilb_df_destroy(ctx->fields.back());      // children die with their context
    while (!ctx->systems.empty()) ilb_particles_destroy(ctx->systems.back());
    if (ctx->gbuffer && ctx->gbuffer_owned) cudaFree(ctx->gbuffer);
    if (ctx->d_lights) cudaFree(ctx->d_lights);
    for (int i = 0; i < 2; i++) {
        if (ctx->h_lights[i]) cudaFreeHost(ctx->h_lights[i]);
        if (ctx->ev_lights[i]) cudaEventDestroy(ctx->ev_lights[i]);
    }
    if (ctx->d_lightmap) cudaFree(ctx->d_lightmap);
    if (ctx->d_probe_in) cudaFree(ctx->d_probe_in);
    if (ctx->d_accum) cudaFree(ctx->d_accum);
    if (ctx->d_accum2) cudaFree(ctx->d_accum2);
    if (ctx->d_tilework) cudaFree(ctx->d_tilework);
    for (ilb_ctx::TileOrder& t : ctx->tile_orders) {
        if (t.h) cudaFreeHost(t.h);
        if (t.d) cudaFree(t.d);
        for (cudaEvent_t e : t.used) if (e) cudaEventDestroy(e);
    }
    if (ctx->light_aux[0]) {
        for (int i = 0; i < 3; i++) { cudaStreamDestroy(ctx->light_aux[i]); cudaEventDestroy(ctx->ev_light_join[i]); }
        cudaEventDestroy(ctx->ev_light_fork);
    }
    if (ctx->d_resolve_in) cudaFree(ctx->d_resolve_in);
    if (ctx->d_resolve_albedo) cudaFree(ctx->d_resolve_albedo);
    if (ctx->d_resolve_out) cudaFree(ctx->d_resolve_out);
    if (ctx->d_resolve_lut) cudaFree(ctx->d_resolve_lut);
    if (ctx->d_ramp_table) cudaFree(ctx->d_ramp_table);
    for (ilb_ctx::RampTexture& t : ctx->ramps) if (t.texels) cudaFree(t.texels);
    for (int i = 0; i < 2; i++) if (ctx->d_luminance[i]) cudaFree(ctx->d_luminance[i]);
    if (ctx->d_plight_scratch) cudaFree(ctx->d_plight_scratch);
    for (ilb_ctx::GBufferUser& u : ctx->gb_users) if (u.done) cudaEventDestroy(u.done);
    for (cudaEvent_t e : ctx->ev_rows_uploaded) if (e) cudaEventDestroy(e);
    if (ctx->band_stream) { cudaStreamDestroy(ctx->band_stream); cudaEventDestroy(ctx->ev_band_fork); cudaEventDestroy(ctx->ev_band_join); }
    if (ctx->copy_in) {
        cudaStreamDestroy(ctx->copy_in);
        cudaStreamDestroy(ctx->copy_out);
        for (int i = 0; i < ILB_PIPELINE_BANDS; i++) { cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_done[i]); }
        for (cudaEvent_t e : ctx->ev_down) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : ctx->ev_frame) if (e) cudaEventDestroy(e);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* ilb_last_error(const ilb_ctx* ctx) {
    if (ctx && live_has(ctx)) return ctx->last_error.c_str();
    std::lock_guard<std::mutex> lock(g_error_mutex);
    static thread_local std::string copy;
    copy = g_create_error;
    return copy.c_str();
}

int ilb_synchronize(ilb_ctx* ctx) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    { const int rcd = ilb_frames_drain(ctx); if (rcd) return rcd; }   // frames in flight finish on the download stream
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_host_register(ilb_ctx* ctx, void* host_ptr, size_t bytes) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!host_ptr || !bytes) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null host range");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { (void)cudaGetLastError(); return ILB_OK; }
    ILB_CUDA(ctx, e);
    return ILB_OK;
}

int ilb_host_unregister(ilb_ctx* ctx, void* host_ptr) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!host_ptr) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null host range");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // nothing of this context may still be copying from / into the range
    const cudaError_t e = cudaHostUnregister(host_ptr);
    if (e == cudaErrorHostMemoryNotRegistered) { (void)cudaGetLastError(); return ILB_OK; }
    ILB_CUDA(ctx, e);
    return ILB_OK;
}

void* ilb_stream(ilb_ctx* ctx) { return (ctx && live_has(ctx)) ? (void*)ctx->stream : nullptr; }
uint64_t ilb_launch_count(const ilb_ctx* ctx) { return (ctx && live_has(ctx)) ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------- distance field
static int df_alloc(ilb_ctx* ctx, int tw, int th, size_t bytes, bool check_bytes, ilb_df** out_df) {
    if (!ctx || !live_has(ctx) || !out_df) return ILB_ERR_INVALID_ARGUMENT;
    *out_df = nullptr;
    if (tw <= 0 || th <= 0 || tw > 8192 || th > 8192)  // DistanceField.MaxSurfaceSize, SDF/DistanceField.cs:19
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "atlas %dx%d outside (0,8192]", tw, th);
    const size_t need = (size_t)8 * tw * th;
    if (check_bytes && bytes != need)  // DistanceField.Load "Truncated file", SDF/DistanceField.cs:200-203
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "atlas needs %zu bytes, got %zu", need, bytes);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ilb_df* df = new ilb_df();
    df->ctx = ctx; df->tw = tw; df->th = th;
    cudaError_t e = cudaMalloc(&df->tex, need);
    if (e != cudaSuccess) {
        delete df;
        return ilb_cuda_fail(ctx, e, "cudaMalloc(distance field)");
    }
    ctx->fields.push_back(df);
    live_add(df);
    *out_df = df;
    return ILB_OK;
}

int ilb_df_create(ilb_ctx* ctx, int tw, int th, const uint16_t* rgba64, size_t bytes, ilb_df** out_df) {
    if (!rgba64) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "rgba64 is null");
    int rc = df_alloc(ctx, tw, th, bytes, true, out_df);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync((*out_df)->tex, rgba64, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        ilb_df_destroy(*out_df);
        *out_df = nullptr;
        return ilb_cuda_fail(ctx, e, "upload distance field");
    }
    return ILB_OK;
}

int ilb_df_create_device(ilb_ctx* ctx, int tw, int th, const void* d_rgba64, size_t bytes, ilb_df** out_df) {
    if (!d_rgba64) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "d_rgba64 is null");
    int rc = df_alloc(ctx, tw, th, bytes, true, out_df);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync((*out_df)->tex, d_rgba64, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        ilb_df_destroy(*out_df);
        *out_df = nullptr;
        return ilb_cuda_fail(ctx, e, "copy distance field");
    }
    return ILB_OK;
}

int ilb_df_generate(ilb_ctx* ctx, int tw, int th, int slice_w, int slice_h, int slice_count, const ilb_df_uniforms* u,
                    const ilb_obstruction* obstructions, int count, ilb_df** out_df) {
    if (!u || count < 0 || (count > 0 && !obstructions)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    int rc = df_alloc(ctx, tw, th, 0, false, out_df);
    if (rc) return rc;
    // slices the atlas has room for but the field does not use stay cleared
    cudaError_t ce = cudaMemsetAsync((*out_df)->tex, 0, (size_t)8 * tw * th, ctx->stream);
    rc = ce != cudaSuccess ? ilb_cuda_fail(ctx, ce, "clear distance field")
                           : ilb_dfgen_launch(ctx, (*out_df)->tex, nullptr, tw, th, slice_w, slice_h, slice_count, u, obstructions, count, nullptr, 0,
                                              nullptr, 0, 0, (slice_count + 2) / 3);
    if (rc) {
        ilb_df_destroy(*out_df);
        *out_df = nullptr;
    }
    return rc;
}

int ilb_df_update_dynamic(ilb_df* df, const ilb_df* static_df, int slice_w, int slice_h, int slice_count, const ilb_df_uniforms* u,
                          const ilb_obstruction* obstructions, int count) {
    if (!df || !live_has(df)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = df->ctx;
    if (!static_df || !live_has(static_df) || static_df->ctx != ctx || static_df == df)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "static field is released, the same as the dynamic field, or belongs to another context");
    if (static_df->tw != df->tw || static_df->th != df->th)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "static field is %dx%d, dynamic field %dx%d", static_df->tw, static_df->th, df->tw, df->th);
    if (!u || count < 0 || (count > 0 && !obstructions)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    df->version++;  // derived planes are rebuilt (in place) the next time the field is sampled
    return ilb_dfgen_launch(ctx, df->tex, static_df->tex, df->tw, df->th, slice_w, slice_h, slice_count, u, obstructions, count, nullptr, 0, nullptr,
                            0, 0, (slice_count + 2) / 3);
}

int ilb_df_create_empty(ilb_ctx* ctx, int tw, int th, ilb_df** out_df) {
    int rc = df_alloc(ctx, tw, th, 0, false, out_df);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemsetAsync((*out_df)->tex, 0, (size_t)8 * tw * th, ctx->stream));
    return ILB_OK;
}

int ilb_df_update_slices(ilb_df* df, const ilb_df* static_df, int slice_w, int slice_h, int slice_count, const ilb_df_uniforms* u,
                         const ilb_obstruction* obstructions, int obstruction_count, const ilb_height_volume* volumes, int volume_count,
                         const ilb_float4* edges, int edge_count, int first_physical_slice, int physical_slice_count) {
    if (!df || !live_has(df)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = df->ctx;
    if (static_df && (!live_has(static_df) || static_df->ctx != ctx || static_df == df || static_df->tw != df->tw || static_df->th != df->th))
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "static field is released, the same as the field, of another size or of another context");
    if (!u || obstruction_count < 0 || (obstruction_count > 0 && !obstructions)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    df->version++;
    return ilb_dfgen_launch(ctx, df->tex, static_df ? static_df->tex : nullptr, df->tw, df->th, slice_w, slice_h, slice_count, u, obstructions,
                            obstruction_count, volumes, volume_count, edges, edge_count, first_physical_slice, physical_slice_count);
}

int ilb_df_download(ilb_df* df, uint16_t* rgba64, size_t bytes) {
    if (!df || !rgba64 || !live_has(df)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = df->ctx;
    const size_t need = (size_t)8 * df->tw * df->th;
    if (bytes != need) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "atlas needs %zu bytes, got %zu", need, bytes);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ILB_CUDA(ctx, cudaMemcpyAsync(rgba64, df->tex, need, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

void ilb_df_destroy(ilb_df* df) {
    if (!df || !live_take(df)) return;
    auto& v = df->ctx->fields;
    v.erase(std::remove(v.begin(), v.end(), df), v.end());
    for (ilb_psys* ps : df->ctx->systems)
        if (ps->field == df) ps->field = nullptr;
    cudaSetDevice(df->ctx->device);
    cudaStreamSynchronize(df->ctx->stream);
    ilb_planes_release(df);
    if (df->tex) cudaFree(df->tex);
    delete df;
}

// ---------------------------------------------------------------------------------------------- G-buffer
static int gbuffer_set(ilb_ctx* ctx, int w, int h, int fmt, const void* data, bool device) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    { const int rcd = ilb_frames_drain(ctx); if (rcd) return rcd; }   // frames in flight read / write the G-buffer
    ctx->gb_generation++;
    if (!data) {  // G-buffer disabled (Configuration.EnableGBuffer == false)
        ctx->gb_w = ctx->gb_h = 0;
        if (ctx->gbuffer && ctx->gbuffer_owned) {
            ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->gbuffer);
        }
        ctx->gbuffer = nullptr;
        ctx->gbuffer_capacity = 0;
        return ILB_OK;
    }
    if (w <= 0 || h <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad G-buffer size %dx%d", w, h);
    if (fmt != ILB_FORMAT_FLOAT4 && fmt != ILB_FORMAT_HALF4)  // GBuffer.cs:31-39
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "G-buffer format must be FLOAT4 or HALF4");
    const size_t bytes = ilb_format_bytes(fmt) * (size_t)w * (size_t)h;
    if (!ctx->gbuffer_owned) { ctx->gbuffer = nullptr; ctx->gbuffer_capacity = 0; }
    int rc = ilb_reserve(ctx, &ctx->gbuffer, &ctx->gbuffer_capacity, bytes, false);
    if (rc) return rc;
    ctx->gbuffer_owned = true;
    ILB_CUDA(ctx, cudaMemcpyAsync(ctx->gbuffer, data, bytes, device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (!device) ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller-owned pageable memory
    ctx->gb_w = w; ctx->gb_h = h; ctx->gb_fmt = fmt;
    if (device) return ilb_gbuffer_note_user(ctx, 0, h);   // a later row upload must not overtake this copy
    return ILB_OK;
}

int ilb_gbuffer_upload_rows(ilb_ctx* ctx, int w, int h, int fmt, int row_begin, int row_end, const void* rows) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!rows || w <= 0 || h <= 0 || row_begin < 0 || row_end > h || row_begin > row_end) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad G-buffer rows [%d,%d) of %dx%d", row_begin, row_end, w, h);
    if (fmt != ILB_FORMAT_FLOAT4 && fmt != ILB_FORMAT_HALF4) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "G-buffer format must be FLOAT4 or HALF4");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    { const int rcd = ilb_frames_drain(ctx); if (rcd) return rcd; }
    ctx->gb_generation++;
    const size_t texel = ilb_format_bytes(fmt), bytes = texel * (size_t)w * (size_t)h;
    if (!ctx->gbuffer_owned) { ctx->gbuffer = nullptr; ctx->gbuffer_capacity = 0; }
    const bool fresh = !ctx->gbuffer || ctx->gb_w != w || ctx->gb_h != h || ctx->gb_fmt != fmt;
    int rc = ilb_reserve(ctx, &ctx->gbuffer, &ctx->gbuffer_capacity, bytes, false);
    if (rc) return rc;
    ctx->gbuffer_owned = true;
    if (fresh) {  // rows that are never uploaded hold zeros
        ILB_CUDA(ctx, cudaMemsetAsync(ctx->gbuffer, 0, bytes, ctx->stream));
        for (ilb_ctx::GBufferUser& u : ctx->gb_users) u.live = false;   // they used the buffer this one replaces
        const int rc2 = ilb_gbuffer_note_user(ctx, 0, h);
        if (rc2) return rc2;
    }
    ctx->gb_w = w; ctx->gb_h = h; ctx->gb_fmt = fmt;
    const size_t off = texel * (size_t)w * (size_t)row_begin, n = texel * (size_t)w * (size_t)(row_end - row_begin);
    if (!n) return ILB_OK;
    // The copy runs on the upload stream, behind exactly the queued work that touches these rows (a band whose kernels are
    // still running does not hold back the upload of the next band), and the context's stream continues behind the copy.
    if (!ctx->copy_in) {
        ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < ILB_PIPELINE_BANDS; i++) {
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        }
    }
    for (ilb_ctx::GBufferUser& u : ctx->gb_users) {
        if (!u.live) continue;
        if (cudaEventQuery(u.done) == cudaSuccess) { u.live = false; continue; }
        if (u.row_begin < row_end && row_begin < u.row_end) ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, u.done, 0));
    }
    ILB_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(ctx->gbuffer) + off, rows, n, cudaMemcpyHostToDevice, ctx->copy_in));
    cudaEvent_t& up = ctx->ev_rows_uploaded[ctx->ev_rows_next];
    ctx->ev_rows_next = (ctx->ev_rows_next + 1) % 8;
    if (!up) ILB_CUDA(ctx, cudaEventCreateWithFlags(&up, cudaEventDisableTiming));
    ILB_CUDA(ctx, cudaEventRecord(up, ctx->copy_in));
    ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, up, 0));
    return ILB_OK;
}

int ilb_gbuffer_upload(ilb_ctx* ctx, int w, int h, int fmt, const void* data) { return gbuffer_set(ctx, w, h, fmt, data, false); }
int ilb_gbuffer_upload_device(ilb_ctx* ctx, int w, int h, int fmt, const void* d) { return gbuffer_set(ctx, w, h, fmt, d, true); }

// ---------------------------------------------------------------------------------------------- lighting
int ilb_render_lighting_device(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches,
                               int batch_count, const ilb_light_vertex* vertices, int vertex_count, void* d_lightmap_out) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    void* outs[1] = {d_lightmap_out};
    return ilb_lighting_launch(ctx, df, frame, batches, batch_count, vertices, vertex_count, outs, 1, false);
}

int ilb_render_lighting_peers(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches,
                              int batch_count, const ilb_light_vertex* vertices, int vertex_count, void* const* d_peer_lightmaps,
                              int peer_count) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!d_peer_lightmaps) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "d_peer_lightmaps is null");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_lighting_launch(ctx, df, frame, batches, batch_count, vertices, vertex_count, d_peer_lightmaps, peer_count, true);
}

int ilb_render_lighting(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches, int batch_count,
                        const ilb_light_vertex* vertices, int vertex_count, void* lightmap_out) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!frame || !lightmap_out) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (frame->width <= 0 || frame->row_end < frame->row_begin) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad frame geometry");
    const size_t bytes = ilb_format_bytes(frame->lightmap_format) * (size_t)frame->width * (size_t)(frame->row_end - frame->row_begin);
    int rc = ilb_reserve(ctx, &ctx->d_lightmap, &ctx->d_lightmap_capacity, std::max<size_t>(bytes, 16), false);
    if (rc) return rc;
    void* outs[1] = {ctx->d_lightmap};
    ctx->lm_fmt = -1;
    rc = ilb_lighting_launch(ctx, df, frame, batches, batch_count, vertices, vertex_count, outs, 1, false);
    if (rc) return rc;
    ctx->lm_w = frame->width; ctx->lm_rows = frame->row_end - frame->row_begin; ctx->lm_fmt = frame->lightmap_format;
    ILB_CUDA(ctx, cudaMemcpyAsync(lightmap_out, ctx->d_lightmap, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_render_lighting_frame(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches, int batch_count,
                              const ilb_light_vertex* vertices, int vertex_count, int gbuffer_width, int gbuffer_height, int gbuffer_format,
                              const void* gbuffer, void* lightmap_out) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->lm_fmt = -1;
    const int rc = ilb_lighting_frame_from_host(ctx, df, frame, batches, batch_count, vertices, vertex_count, gbuffer_width, gbuffer_height,
                                                gbuffer_format, gbuffer, lightmap_out, nullptr);
    if (rc == ILB_OK) { ctx->lm_w = frame->width; ctx->lm_rows = frame->row_end - frame->row_begin; ctx->lm_fmt = frame->lightmap_format; }
    return rc;
}

int ilb_render_lighting_frame_async(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches, int batch_count,
                                    const ilb_light_vertex* vertices, int vertex_count, int gbuffer_width, int gbuffer_height, int gbuffer_format,
                                    const void* gbuffer, void* lightmap_out, uint64_t* out_ticket) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!out_ticket) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null ticket");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->lm_fmt = -1;
    unsigned long long ticket = 0;
    const int rc = ilb_lighting_frame_from_host(ctx, df, frame, batches, batch_count, vertices, vertex_count, gbuffer_width, gbuffer_height,
                                                gbuffer_format, gbuffer, lightmap_out, &ticket);
    *out_ticket = ticket;
    if (rc == ILB_OK) { ctx->lm_w = frame->width; ctx->lm_rows = frame->row_end - frame->row_begin; ctx->lm_fmt = frame->lightmap_format; }
    return rc;
}

int ilb_render_lighting_frame_wait(ilb_ctx* ctx, uint64_t ticket) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_lighting_frame_wait(ctx, ticket);
}

// ---------------------------------------------------------------------------------------------- resolve / luminance (N3)
// The lightmap a resolve / luminance call reads: the caller's host texels (staged), or the context's resident one.
static int resolve_source(ilb_ctx* ctx, int w, int h, int fmt, const void* lightmap_host, const void** d_out) {
    if (w <= 0 || h <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad size %dx%d", w, h);
    if (fmt != ILB_FORMAT_FLOAT4 && fmt != ILB_FORMAT_HALF4 && fmt != ILB_FORMAT_RGBA8) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad lightmap format %d", fmt);
    if (!lightmap_host) {
        { const int rcd = ilb_frames_drain(ctx); if (rcd) return rcd; }   // the resident lightmap is the last frame's, complete
        if (ctx->lm_fmt < 0 || !ctx->d_lightmap) return ilb_fail(ctx, ILB_ERR_INVALID_OPERATION, "no resident lightmap: render a frame first or pass the texels");
        if (ctx->lm_w != w || ctx->lm_rows != h || ctx->lm_fmt != fmt)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "resident lightmap is %dx%d format %d, asked for %dx%d format %d", ctx->lm_w, ctx->lm_rows, ctx->lm_fmt, w, h, fmt);
        *d_out = ctx->d_lightmap;
        return ILB_OK;
    }
    const size_t bytes = ilb_format_bytes(fmt) * (size_t)w * (size_t)h;
    int rc = ilb_reserve(ctx, &ctx->d_resolve_in, &ctx->d_resolve_in_capacity, bytes, false);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_resolve_in, lightmap_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *d_out = ctx->d_resolve_in;
    return ILB_OK;
}

int ilb_resolve_lighting_device(ilb_ctx* ctx, const ilb_resolve* params, const void* d_lightmap, const void* d_albedo, void* d_output) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!params || !d_lightmap || !d_output) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_resolve_launch(ctx, params, d_lightmap, d_albedo, d_output);
}

int ilb_resolve_lighting(ilb_ctx* ctx, const ilb_resolve* params, const void* lightmap, const void* albedo, void* output) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!params || !output) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_lm = nullptr;
    int rc = resolve_source(ctx, params->width, params->height, params->lightmap_format, lightmap, &d_lm);
    if (rc) return rc;
    const size_t n = (size_t)params->width * (size_t)params->height;
    const void* d_al = nullptr;
    if (albedo) {
        if (params->albedo_format != ILB_FORMAT_FLOAT4 && params->albedo_format != ILB_FORMAT_RGBA8)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "albedo format must be RGBA8 or FLOAT4");
        const size_t abytes = ilb_format_bytes(params->albedo_format) * n;
        rc = ilb_reserve(ctx, &ctx->d_resolve_albedo, &ctx->d_resolve_albedo_capacity, abytes, false);
        if (rc) return rc;
        ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_resolve_albedo, albedo, abytes, cudaMemcpyHostToDevice, ctx->stream));
        d_al = ctx->d_resolve_albedo;
    }
    if (params->output_format != ILB_FORMAT_FLOAT4 && params->output_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "output format must be RGBA8 or FLOAT4");
    const size_t obytes = ilb_format_bytes(params->output_format) * n;
    rc = ilb_reserve(ctx, &ctx->d_resolve_out, &ctx->d_resolve_out_capacity, obytes, false);
    if (rc) return rc;
    rc = ilb_resolve_launch(ctx, params, d_lm, d_al, ctx->d_resolve_out);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemcpyAsync(output, ctx->d_resolve_out, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

// ---------------------------------------------------------------------------------------------- ramp textures
int ilb_ramp_texture_create(ilb_ctx* ctx, int width, int height, int format, const void* texels, int32_t* out_id) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!texels || !out_id || width <= 0 || height <= 0 || (size_t)width * (size_t)height > ((size_t)1 << 26))
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad ramp texture %dx%d", width, height);
    if (format != ILB_FORMAT_RGBA8 && format != ILB_FORMAT_FLOAT4) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "ramp texture format must be RGBA8 or FLOAT4");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)width * (size_t)height;
    std::vector<float> decoded(4 * n);   // the kernel samples float4 texels; UNORM8 -> float is c / 255
    if (format == ILB_FORMAT_FLOAT4) memcpy(decoded.data(), texels, sizeof(float) * 4 * n);
    else for (size_t i = 0; i < 4 * n; i++) decoded[i] = (float)reinterpret_cast<const unsigned char*>(texels)[i] / 255.0f;
    ilb_ctx::RampTexture t;
    t.w = width; t.h = height;
    ILB_CUDA(ctx, cudaMalloc(&t.texels, sizeof(float4) * n));
    const cudaError_t e = cudaMemcpy(t.texels, decoded.data(), sizeof(float4) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(t.texels); return ilb_cuda_fail(ctx, e, "upload a ramp texture"); }
    size_t slot = ctx->ramps.size();
    for (size_t i = 0; i < ctx->ramps.size(); i++) if (!ctx->ramps[i].texels) { slot = i; break; }   // reuse a destroyed id
    if (slot == ctx->ramps.size()) ctx->ramps.push_back(t); else ctx->ramps[slot] = t;
    ctx->ramp_table_dirty = true;
    *out_id = (int32_t)slot + 1;
    return ILB_OK;
}

int ilb_ramp_texture_destroy(ilb_ctx* ctx, int32_t id) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (id < 1 || (size_t)id > ctx->ramps.size() || !ctx->ramps[(size_t)id - 1].texels) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "unknown ramp texture %d", id);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // a frame in flight may still sample it
    cudaFree(ctx->ramps[(size_t)id - 1].texels);
    ctx->ramps[(size_t)id - 1] = ilb_ctx::RampTexture();
    ctx->ramp_table_dirty = true;
    return ILB_OK;
}

int ilb_set_dithering(ilb_ctx* ctx, const ilb_dithering* settings) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    const ilb_dithering defaults = {0.0f, 255.0f, 0.0f, 1.0f, 0.0f, 1.0f};
    ctx->dither = settings ? *settings : defaults;
    if (!(ctx->dither.Unit > 0.0f) && ctx->dither.Unit != 0.0f) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "Dithering.Unit must be positive");
    return ILB_OK;
}

int ilb_resolve_lighting_lut_device(ilb_ctx* ctx, const ilb_resolve* params, const ilb_lut_blending* lut, const void* d_dark_lut,
                                    const void* d_bright_lut, const void* d_lightmap, const void* d_albedo, void* d_output) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!params || !lut || !d_dark_lut || !d_bright_lut || !d_lightmap || !d_output) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (!d_albedo) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "LUT blending is not compatible with this type of lighting resolve (no albedo)");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_resolve_lut_launch(ctx, params, lut, d_dark_lut, d_bright_lut, d_lightmap, d_albedo, d_output);
}

int ilb_resolve_lighting_lut(ilb_ctx* ctx, const ilb_resolve* params, const ilb_lut_blending* lut, const void* dark_lut, const void* bright_lut,
                             const void* lightmap, const void* albedo, void* output) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!params || !lut || !dark_lut || !bright_lut || !output) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (!albedo) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "LUT blending is not compatible with this type of lighting resolve (no albedo)");
    if (lut->dark_resolution < 2 || lut->bright_resolution < 2 || lut->dark_row_count < 1 || lut->bright_row_count < 1 ||
        lut->dark_resolution > 256 || lut->bright_resolution > 256 || lut->dark_row_count > 256 || lut->bright_row_count > 256)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad LUT geometry");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_lm = nullptr;
    int rc = resolve_source(ctx, params->width, params->height, params->lightmap_format, lightmap, &d_lm);
    if (rc) return rc;
    if (params->albedo_format != ILB_FORMAT_FLOAT4 && params->albedo_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "albedo format must be RGBA8 or FLOAT4");
    if (params->output_format != ILB_FORMAT_FLOAT4 && params->output_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "output format must be RGBA8 or FLOAT4");
    const size_t n = (size_t)params->width * (size_t)params->height;
    const size_t abytes = ilb_format_bytes(params->albedo_format) * n, obytes = ilb_format_bytes(params->output_format) * n;
    auto lutBytes = [](int res, int rows) { return (size_t)4 * (size_t)res * res * (size_t)res * rows; };
    const size_t dbytes = (lutBytes(lut->dark_resolution, lut->dark_row_count) + 255) & ~(size_t)255, bbytes = lutBytes(lut->bright_resolution, lut->bright_row_count);
    rc = ilb_reserve(ctx, &ctx->d_resolve_albedo, &ctx->d_resolve_albedo_capacity, abytes, false);
    if (rc) return rc;
    rc = ilb_reserve(ctx, &ctx->d_resolve_out, &ctx->d_resolve_out_capacity, obytes, false);
    if (rc) return rc;
    rc = ilb_reserve(ctx, &ctx->d_resolve_lut, &ctx->d_resolve_lut_capacity, dbytes + bbytes, false);
    if (rc) return rc;
    char* d_lut = reinterpret_cast<char*>(ctx->d_resolve_lut);
    ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_resolve_albedo, albedo, abytes, cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaMemcpyAsync(d_lut, dark_lut, lutBytes(lut->dark_resolution, lut->dark_row_count), cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaMemcpyAsync(d_lut + dbytes, bright_lut, bbytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = ilb_resolve_lut_launch(ctx, params, lut, d_lut, d_lut + dbytes, d_lm, ctx->d_resolve_albedo, ctx->d_resolve_out);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemcpyAsync(output, ctx->d_resolve_out, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_resolve_lighting_placed_device(ilb_ctx* ctx, const ilb_resolve* params, const ilb_resolve_placement* placement, const void* d_lightmap,
                                       const void* d_albedo, void* d_target) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!params || !placement || !d_lightmap || !d_target) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_resolve_placed_launch(ctx, params, placement, d_lightmap, d_albedo, d_target);
}

int ilb_resolve_lighting_placed(ilb_ctx* ctx, const ilb_resolve* params, const ilb_resolve_placement* pl, const void* lightmap, const void* albedo,
                                void* target) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!params || !pl || !target) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (pl->target_width <= 0 || pl->target_height <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad target size");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_lm = nullptr;
    int rc = resolve_source(ctx, params->width, params->height, params->lightmap_format, lightmap, &d_lm);
    if (rc) return rc;
    const void* d_al = nullptr;
    if (albedo) {
        if (params->albedo_format != ILB_FORMAT_FLOAT4 && params->albedo_format != ILB_FORMAT_RGBA8)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "albedo format must be RGBA8 or FLOAT4");
        if (pl->albedo_width <= 0 || pl->albedo_height <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad albedo size");
        const size_t abytes = ilb_format_bytes(params->albedo_format) * (size_t)pl->albedo_width * (size_t)pl->albedo_height;
        rc = ilb_reserve(ctx, &ctx->d_resolve_albedo, &ctx->d_resolve_albedo_capacity, abytes, false);
        if (rc) return rc;
        ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_resolve_albedo, albedo, abytes, cudaMemcpyHostToDevice, ctx->stream));
        d_al = ctx->d_resolve_albedo;
    }
    if (params->output_format != ILB_FORMAT_FLOAT4 && params->output_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "output format must be RGBA8 or FLOAT4");
    const size_t obytes = ilb_format_bytes(params->output_format) * (size_t)pl->target_width * (size_t)pl->target_height;
    rc = ilb_reserve(ctx, &ctx->d_resolve_out, &ctx->d_resolve_out_capacity, obytes, false);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_resolve_out, target, obytes, cudaMemcpyHostToDevice, ctx->stream));   // pixels outside the quad are kept
    rc = ilb_resolve_placed_launch(ctx, params, pl, d_lm, d_al, ctx->d_resolve_out);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemcpyAsync(target, ctx->d_resolve_out, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_compute_luminance(ilb_ctx* ctx, int width, int height, int lightmap_format, const void* lightmap, int level, float* out_luminance) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!out_luminance) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_lm = nullptr;
    int rc = resolve_source(ctx, width, height, lightmap_format, lightmap, &d_lm);
    if (rc) return rc;
    return ilb_luminance_launch(ctx, width, height, lightmap_format, d_lm, level, out_luminance);
}

int ilb_lighting_set_particle_lights(ilb_ctx* ctx, const ilb_particle_light_source* sources, int count) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (count < 0 || (count > 0 && !sources)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null or negative argument");
    for (int i = 0; i < count; i++)
        if (!sources[i].system || !live_has(sources[i].system) || sources[i].system->ctx != ctx)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "particle light source %d: system is released or belongs to another context", i);
    ctx->particle_lights.assign(sources, sources + count);
    return ILB_OK;
}

int ilb_update_light_probes(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches, int batch_count,
                            const ilb_light_vertex* vertices, int vertex_count, const ilb_float4* probe_positions,
                            const ilb_float4* probe_normals, int probe_count, int output_format, void* probes_out) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_probes_launch(ctx, df, frame, batches, batch_count, vertices, vertex_count, probe_positions, probe_normals, probe_count,
                             output_format, probes_out, nullptr);
}

int ilb_update_light_probes_device(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* frame, const ilb_light_batch* batches, int batch_count,
                                   const ilb_light_vertex* vertices, int vertex_count, const ilb_float4* probe_positions,
                                   const ilb_float4* probe_normals, int probe_count, int output_format, void* d_probes_out) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    if (!d_probes_out) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "d_probes_out is null");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_probes_launch(ctx, df, frame, batches, batch_count, vertices, vertex_count, probe_positions, probe_normals, probe_count,
                             output_format, nullptr, d_probes_out);
}

// ---------------------------------------------------------------------------------------------- particles
int ilb_particles_create(ilb_ctx* ctx, int chunk_size, int max_chunks, ilb_psys** out_psys) {
    if (!ctx || !live_has(ctx) || !out_psys) return ILB_ERR_INVALID_ARGUMENT;
    *out_psys = nullptr;
    if (chunk_size < 16 || chunk_size > 4096 || max_chunks < 1)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "chunk_size %d / max_chunks %d out of range", chunk_size, max_chunks);
    const size_t per = (size_t)chunk_size * chunk_size, total = per * (size_t)max_chunks;
    if (total >= (size_t)1 << 31) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "more than 2^31 particles in one system");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ilb_psys* ps = new ilb_psys();
    ps->ctx = ctx; ps->chunk_size = chunk_size; ps->max_chunks = max_chunks; ps->per_chunk = per;
    ctx->systems.push_back(ps);
    live_add(ps);
    if (const char* e = getenv("ILB_PARTICLE_TMA")) ps->use_tma = (e[0] != '0');
    cudaDeviceGetAttribute(&ps->sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 5 && e == cudaSuccess; i++) {
        e = cudaMalloc(&ps->buf[i], sizeof(float4) * total);
        if (e == cudaSuccess) e = cudaMemsetAsync(ps->buf[i], 0, sizeof(float4) * total, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMalloc(&ps->d_count, sizeof(unsigned long long));
    if (e != cudaSuccess) {
        ilb_particles_destroy(ps);
        return ilb_cuda_fail(ctx, e, "allocate particle storage");
    }
    *out_psys = ps;
    return ILB_OK;
}

void ilb_particles_destroy(ilb_psys* ps) {
    if (!ps || !live_take(ps)) return;
    auto& v = ps->ctx->systems;
    v.erase(std::remove(v.begin(), v.end(), ps), v.end());
    auto& pl = ps->ctx->particle_lights;
    pl.erase(std::remove_if(pl.begin(), pl.end(), [ps](const ilb_particle_light_source& s) { return s.system == ps; }), pl.end());
    cudaSetDevice(ps->ctx->device);
    cudaStreamSynchronize(ps->ctx->stream);
    for (int i = 0; i < 5; i++)
        if (ps->buf[i]) cudaFree(ps->buf[i]);
    if (ps->rng) cudaFree(ps->rng);
    if (ps->noise_table) cudaFree(ps->noise_table);
    if (ps->escape_table) cudaFree(ps->escape_table);
    if (ps->positions) cudaFree(ps->positions);
    if (ps->pattern) cudaFree(ps->pattern);
    ilb_raster_release(ps);
    if (ps->life_ramp) cudaFree(ps->life_ramp);
    if (ps->d_count) cudaFree(ps->d_count);
    if (ps->d_chunk_counts) cudaFree(ps->d_chunk_counts);
    if (ps->h_chunk_counts) cudaFreeHost(ps->h_chunk_counts);
    if (ps->ev_chunk_counts) cudaEventDestroy(ps->ev_chunk_counts);
    delete ps;
}

int ilb_particles_set_randomness(ilb_psys* ps, const ilb_float4* table, int w, int h) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = ps->ctx;
    if (!table || w <= 0 || h <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad randomness table");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ps->rng) cudaFree(ps->rng);
    ps->rng = nullptr;
    const size_t bytes = sizeof(float4) * (size_t)w * (size_t)h;
    ILB_CUDA(ctx, cudaMalloc(&ps->rng, bytes));
    ILB_CUDA(ctx, cudaMemcpyAsync(ps->rng, table, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ps->rng_w = w; ps->rng_h = h;
    return ILB_OK;
}

int ilb_particles_set_life_ramp(ilb_psys* ps, const ilb_float4* texels, int w, int h) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = ps->ctx;
    if (texels && (w <= 0 || h <= 0 || w > 16384 || h > 16384)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad life ramp size %dx%d", w, h);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ps->life_ramp) cudaFree(ps->life_ramp);
    ps->life_ramp = nullptr;
    ps->life_ramp_w = ps->life_ramp_h = 0;
    if (!texels) return ILB_OK;
    const size_t bytes = sizeof(float4) * (size_t)w * (size_t)h;
    ILB_CUDA(ctx, cudaMalloc(&ps->life_ramp, bytes));
    ILB_CUDA(ctx, cudaMemcpyAsync(ps->life_ramp, texels, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ps->life_ramp_w = w; ps->life_ramp_h = h;
    return ILB_OK;
}

int ilb_particles_set_collision_field(ilb_psys* ps, ilb_df* df) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    if (df && (!live_has(df) || df->ctx != ps->ctx)) return ilb_fail(ps->ctx, ILB_ERR_INVALID_ARGUMENT, "distance field is released or belongs to another context");
    ps->field = df;
    return ILB_OK;
}

int ilb_particles_set_live_chunks(ilb_psys* ps, int count) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    if (count < 0 || count > ps->max_chunks) return ilb_fail(ps->ctx, ILB_ERR_INVALID_ARGUMENT, "live chunk count %d outside [0,%d]", count, ps->max_chunks);
    ps->live_chunks = count;
    return ILB_OK;
}

int ilb_particles_upload_chunk(ilb_psys* ps, int chunk, const ilb_float4* p, const ilb_float4* v, const ilb_float4* a) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = ps->ctx;
    if (chunk < 0 || chunk >= ps->max_chunks) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "chunk %d outside [0,%d)", chunk, ps->max_chunks);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(float4) * ps->per_chunk, off = ps->per_chunk * (size_t)chunk;
    const ilb_float4* src[3] = {p, v, a};
    for (int i = 0; i < 3; i++)
        if (src[i]) ILB_CUDA(ctx, cudaMemcpyAsync(ps->buf[i] + off, src[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_particles_upload_buffer(ilb_psys* ps, int chunk, int which, const ilb_float4* data) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = ps->ctx;
    if (!data || which < 0 || which > 4) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null data or bad buffer index %d", which);
    if (chunk < 0 || chunk >= ps->max_chunks) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "chunk %d out of range [0,%d)", chunk, ps->max_chunks);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    ILB_CUDA(ctx, cudaMemcpyAsync(ps->buf[which] + (size_t)chunk * ps->per_chunk, data, sizeof(float4) * ps->per_chunk, cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_particles_download_chunk(ilb_psys* ps, int chunk, ilb_float4* p, ilb_float4* v, ilb_float4* a, ilb_float4* rc, ilb_float4* rd) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = ps->ctx;
    if (chunk < 0 || chunk >= ps->max_chunks) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "chunk %d outside [0,%d)", chunk, ps->max_chunks);
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(float4) * ps->per_chunk, off = ps->per_chunk * (size_t)chunk;
    ilb_float4* dst[5] = {p, v, a, rc, rd};
    for (int i = 0; i < 5; i++)
        if (dst[i]) ILB_CUDA(ctx, cudaMemcpyAsync(dst[i], ps->buf[i] + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_particles_step(ilb_psys* ps, const ilb_psys_uniforms* u, const ilb_spawn* spawns, int spawn_count, const ilb_op* ops,
                       int op_count, int steps) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    return ilb_particles_launch(ps, u, spawns, nullptr, spawn_count, ops, op_count, steps);
}

int ilb_particles_step_sources(ilb_psys* ps, const ilb_psys_uniforms* u, const ilb_spawn* spawns, const ilb_spawn_source* sources,
                               int spawn_count, const ilb_op* ops, int op_count, int steps) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    if (sources)
        for (int i = 0; i < spawn_count; i++)
            if (sources[i].kind == ILB_SPAWN_FEEDBACK && (!sources[i].source_system || !live_has(sources[i].source_system)))
                return ilb_fail(ps->ctx, ILB_ERR_INVALID_ARGUMENT, "spawn %d: the feedback source system is null or released", i);
    return ilb_particles_launch(ps, u, spawns, sources, spawn_count, ops, op_count, steps);
}

// ---------------------------------------------------------------------------------------------- particle rasterisation (N2)
int ilb_particles_render_device(ilb_psys* ps, const ilb_particle_render* params, const void* d_texture, void* d_target) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    return ilb_raster_launch(ps, params, d_texture, d_target);
}

int ilb_particles_render(ilb_psys* ps, const ilb_particle_render* r, const void* texture, void* target) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ilb_ctx* ctx = ps->ctx;
    if (!r || !target) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (r->width <= 0 || r->height <= 0 || r->width > 32768 || r->height > 32768) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad target size %dx%d", r->width, r->height);
    if (r->target_format != ILB_FORMAT_FLOAT4 && r->target_format != ILB_FORMAT_HALF4 && r->target_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad target format %d", r->target_format);
    const void* d_tex = nullptr;
    if (r->texture_filter != ILB_TEXTURE_NONE) {
        if (!texture || r->texture_width < 1 || r->texture_height < 1 || r->texture_width > 16384 || r->texture_height > 16384)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "textured material without a texture");
        const size_t tbytes = (size_t)r->texture_width * (size_t)r->texture_height * 4;
        int rc = ilb_reserve(ctx, &ps->raster[9], &ps->raster_capacity[9], tbytes, false);
        if (rc) return rc;
        ILB_CUDA(ctx, cudaMemcpyAsync(ps->raster[9], texture, tbytes, cudaMemcpyHostToDevice, ctx->stream));
        d_tex = ps->raster[9];
    }
    const size_t bytes = ilb_format_bytes(r->target_format) * (size_t)r->width * (size_t)r->height;
    int rc = ilb_reserve(ctx, &ps->raster[10], &ps->raster_capacity[10], bytes, false);
    if (rc) return rc;
    if (!r->clear) ILB_CUDA(ctx, cudaMemcpyAsync(ps->raster[10], target, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = ilb_raster_launch(ps, r, d_tex, ps->raster[10]);
    if (rc) return rc;
    ILB_CUDA(ctx, cudaMemcpyAsync(target, ps->raster[10], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}

int ilb_particles_composite_layers(ilb_ctx* ctx, const void* const* d_layers, int layer_count, int width, int height, int row_begin,
                                   int row_end, int blend, int target_format, const ilb_float4* clear_color, void* const* d_targets,
                                   int target_count) {
    if (!ctx || !live_has(ctx)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ilb_raster_composite(ctx, d_layers, layer_count, width, height, row_begin, row_end, blend, target_format, clear_color, d_targets,
                                target_count);
}

void* ilb_particles_device_buffer(ilb_psys* ps, int which) {
    if (!ps || !live_has(ps) || which < 0 || which > 4) return nullptr;
    return ps->buf[which];
}

int ilb_particles_count_live(ilb_psys* ps, int64_t* out_count) {
    if (!ps || !out_count || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    return ilb_particles_count_launch(ps, out_count);
}

int ilb_particles_request_chunk_liveness(ilb_psys* ps) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    return ilb_particles_liveness_request(ps);
}

int ilb_particles_poll_chunk_liveness(ilb_psys* ps, int64_t* counts, int capacity, int* out_count, int wait) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    if (!counts || !out_count || capacity < 0) return ilb_fail(ps->ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    return ilb_particles_liveness_poll(ps, counts, capacity, out_count, wait);
}

int ilb_particles_remove_chunk(ilb_psys* ps, int chunk) {
    if (!ps || !live_has(ps)) return ILB_ERR_INVALID_ARGUMENT;
    ILB_CUDA(ps->ctx, cudaSetDevice(ps->ctx->device));
    return ilb_particles_remove(ps, chunk);
}

}  // extern "C"
