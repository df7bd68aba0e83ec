// Expanded distance-field planes (see sampleFieldPlanesT in ilb_device.cuh): built once per (atlas, addressing
// uniforms) on the device and cached on the ilb_df handle.  The atlas stays the interchange format
// (DistanceField.Save / Load, SDF/DistanceField.cs:178-213); the planes are a derived acceleration structure that
// trades HBM capacity (16 B per texel per virtual slice; C4: 1.3 GB of the 180 GB) for per-sample instructions in the
// cone trace, which is issue-bound, not bandwidth-bound.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ilb_internal.h"

namespace {

constexpr int HALO = 2;

struct PlaneBuildParams {
    const uint2* tex;
    float4* planes;
    int tw, th, sw, sh, pw, ph, nv;
    int columnsForWrap;  // unused by the kernel (wrap is modulo tw); kept for debugging
    int col[ILB_MAX_VIRTUAL_SLICES], row[ILB_MAX_VIRTUAL_SLICES];
};

__global__ void __launch_bounds__(256) df_planes_build_kernel(const __grid_constant__ PlaneBuildParams P) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y, v = blockIdx.z;
    if (px >= P.pw) return;
    const int lx = px - HALO, ly = py - HALO;
    // the atlas sampler's addressing: U wraps modulo the atlas width, V clamps (DistanceFieldCommon.fxh:273-281)
    int ax = (P.col[v] * P.sw + lx) % P.tw;
    if (ax < 0) ax += P.tw;
    int ax1 = ax + 1;
    if (ax1 == P.tw) ax1 = 0;
    const int ay = min(max(P.row[v] * P.sh + ly, 0), P.th - 1);
    const uint2 t0 = __ldg(P.tex + (size_t)ay * P.tw + ax), t1 = __ldg(P.tex + (size_t)ay * P.tw + ax1);
    const uint32_t sel = 0x3210u + 0x2222u * (uint32_t)(v % 3);
    const uint32_t p0 = __byte_perm(t0.x, t0.y, sel), p1 = __byte_perm(t1.x, t1.y, sel);
    const float k = 1.0f / 65535.0f;
    const float lo0 = xmul(u16lo(p0), k), hi0 = xmul(u16hi(p0), k), lo1 = xmul(u16lo(p1), k), hi1 = xmul(u16hi(p1), k);
    P.planes[((size_t)v * P.ph + py) * P.pw + px] = make_float4(lo0, hi0, xsub(lo1, lo0), xsub(hi1, hi0));
}

// TMA-staged form: the three virtual slices 3c, 3c + 1, 3c + 2 read the SAME atlas texels (channel pairs (r, g), (g, b),
// (b, a) of physical slice c), and every entry needs its right-hand neighbour as well, so each atlas texel feeds six plane
// entries.  One elected thread stages the CTA's run of an atlas row (257 texels, about 2 KB) into shared memory with a single
// bulk asynchronous copy (cp.async.bulk, completion on an mbarrier; SASS: UBLKCP) and the 256 threads emit up to three 16-byte
// entries each from it: the atlas is read once and only by the copy engine, the threads issue stores only.  CTAs whose run
// crosses the U wrap of the atlas, or whose 16-byte aligned run would leave the allocation, take the direct loads instead
// (a few columns at the atlas edges) -- both forms run the same arithmetic, so the planes are bit-identical.
constexpr int BUILD_THREADS = 256;
__global__ void __launch_bounds__(BUILD_THREADS) df_planes_build_tma_kernel(const __grid_constant__ PlaneBuildParams P) {
    __shared__ __align__(16) uint2 s_row[BUILD_THREADS + 4];
    __shared__ unsigned long long s_bar;
    const int tid = threadIdx.x, px = blockIdx.x * BUILD_THREADS + tid, py = blockIdx.y, c = blockIdx.z;  // c: physical slice
    const int v0 = 3 * c;
    const int ay = min(max(P.row[v0] * P.sh + (py - HALO), 0), P.th - 1);            // V clamps
    // atlas x of this CTA's first entry; U wraps modulo the atlas width (that is also how a slice index reaches the cells of the
    // atlas's second and later rows: col = v / 3 runs past the column count, DistanceFieldCommon.fxh:273-281, :327-337)
    long long first = ((long long)P.col[v0] * P.sw + (long long)blockIdx.x * BUILD_THREADS - HALO) % P.tw;
    if (first < 0) first += P.tw;
    const int count = min(BUILD_THREADS, P.pw - blockIdx.x * BUILD_THREADS) + 1;     // texels the CTA reads: its entries + one neighbour
    const long long g0 = (long long)ay * P.tw + first, ga = g0 & ~1LL;               // 16-byte aligned start (the atlas is 256-byte aligned)
    const long long gend = (g0 + count + 1) & ~1LL;
    const bool staged = (first + count <= P.tw) && (gend <= (long long)P.tw * P.th); // uniform over the CTA: the run does not wrap
    uint2 t0, t1;
    if (staged) {
        if (tid == 0) {
            mbarInit(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const unsigned bytes = (unsigned)(gend - ga) * (unsigned)sizeof(uint2);
            mbarExpectTx(&s_bar, bytes);
            bulkLoad(s_row, P.tex + ga, bytes, &s_bar);
        }
        __syncthreads();          // the barrier is initialised before anyone waits on it
        mbarWait(&s_bar, 0);
        const int at = (int)(g0 - ga) + tid;
        if (px < P.pw) { t0 = s_row[at]; t1 = s_row[at + 1]; }
    } else if (px < P.pw) {       // the run crosses the wrap: per-thread loads
        const int ax = (int)((first + tid) % P.tw);
        int ax1 = ax + 1;
        if (ax1 == P.tw) ax1 = 0;
        t0 = __ldg(P.tex + (size_t)ay * P.tw + ax);
        t1 = __ldg(P.tex + (size_t)ay * P.tw + ax1);
    }
    if (px >= P.pw) return;
    const float k = 1.0f / 65535.0f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int v = v0 + j;
        if (v >= P.nv) break;
        const uint32_t sel = 0x3210u + 0x2222u * (uint32_t)j;
        const uint32_t p0 = __byte_perm(t0.x, t0.y, sel), p1 = __byte_perm(t1.x, t1.y, sel);
        const float lo0 = xmul(u16lo(p0), k), hi0 = xmul(u16hi(p0), k), lo1 = xmul(u16lo(p1), k), hi1 = xmul(u16hi(p1), k);
        __stcs(P.planes + ((size_t)v * P.ph + py) * P.pw + px, make_float4(lo0, hi0, xsub(lo1, lo0), xsub(hi1, hi0)));
    }
}

// ILB_PLANES_TMA=0 selects the one-thread-per-entry kernel (A/B and the bit-identity test)
void launchPlaneBuild(ilb_ctx* ctx, const PlaneBuildParams& B) {
    bool tma = true;
    if (const char* e = getenv("ILB_PLANES_TMA")) tma = e[0] != '0';
    // the staged kernel assumes what the host arithmetic of ilb_planes_attach guarantees for regular atlases: the three virtual
    // slices of a physical slice share their cell
    for (int v = 0; v < B.nv && tma; v++) tma = (B.col[v] == B.col[3 * (v / 3)]) && (B.row[v] == B.row[3 * (v / 3)]);
    if (tma) df_planes_build_tma_kernel<<<dim3((B.pw + BUILD_THREADS - 1) / BUILD_THREADS, B.ph, (B.nv + 2) / 3), BUILD_THREADS, 0, ctx->stream>>>(B);
    else df_planes_build_kernel<<<dim3((B.pw + 255) / 256, B.ph, B.nv), 256, 0, ctx->stream>>>(B);
    ctx->launches++;
}

// the slice table as floorBiased() indexes it: entry (ILB_FLOOR_BIAS + v) of the returned pointer is entry v of the table
const float4* biasedTable(const float4* vtab) {
    return reinterpret_cast<const float4*>(reinterpret_cast<uintptr_t>(vtab) - (uintptr_t)ILB_FLOOR_BIAS * sizeof(float4));
}

bool sameKey(const ilb_df_planes& p, const DFGeometry& g) {
    return p.key[0] == g.sliceSizeX && p.key[1] == g.sliceSizeY && p.key[2] == g.texelSizeX && p.key[3] == g.texelSizeY &&
           p.key[4] == g.ex && p.key[5] == g.ey && p.key[6] == g.maxValidZ && p.key[7] == g.zToSlice &&
           p.key[8] == g.invSliceCountXTimesOneThird && p.key[9] == g.sliceCount;
}

}  // namespace

void ilb_planes_release(ilb_df* df) {
    for (ilb_df_planes& p : df->planes) {
        if (p.planes) cudaFree(p.planes);
        if (p.vtab) cudaFree(p.vtab);
    }
    df->planes.clear();
}

// Fills g->planes / vtab / pitch when the planes layout applies to these uniforms; leaves them null (atlas sampler)
// when the uniforms do not describe a regular column x row atlas, when an index could leave the halo, or when the
// allocation does not fit.  `columns` / `rows` are TextureSliceCount.xy.
int ilb_planes_attach(ilb_ctx* ctx, ilb_df* df, const ilb_df_uniforms& u, DFGeometry* g) {
    g->planes = nullptr; g->vtab = nullptr; g->vtabBiased = nullptr; g->pitch = 0;
    if (!df || !g->tex) return ILB_OK;
    if (const char* e = getenv("ILB_NO_PLANES"))  // read per call: the parity tests flip it to compare both samplers
        if (e[0] != '0') return ILB_OK;
    ilb_df_planes* stale = nullptr;
    for (ilb_df_planes& p : df->planes)
        if (sameKey(p, *g)) {
            if (p.version == df->version) {
                g->planes = p.planes; g->vtab = p.vtab; g->vtabBiased = biasedTable(p.vtab); g->pitch = p.pitch;
                return ILB_OK;
            }
            stale = &p;  // the atlas was rewritten in place: same geometry, same allocation, new contents
        }
    const int columns = (int)u.TextureSliceCount.x, rows = (int)u.TextureSliceCount.y;
    if (columns < 1 || rows < 1 || (float)columns != u.TextureSliceCount.x || (float)rows != u.TextureSliceCount.y) return ILB_OK;
    if (df->tw % columns || df->th % rows) return ILB_OK;
    const int sw = df->tw / columns, sh = df->th / rows;
    // number of virtual slice indices the sampler can produce: floor(min(z, maxValidZ) * zToSlice), z >= 0
    const float top = g->maxValidZ * g->zToSlice;
    if (!(top >= 0.0f) || !(top < (float)ILB_MAX_VIRTUAL_SLICES)) return ILB_OK;
    if (!(g->maxValidZ >= 0.0f) || !(g->zToSlice >= 0.0f)) return ILB_OK;
    const int nv = (int)std::floor(top) + 1;
    const int pw = sw + 2 * HALO, ph = sh + 2 * HALO;
    if ((double)nv * pw * ph >= 2.0e9) return ILB_OK;  // 32-bit entry indices

    PlaneBuildParams B;
    memset(&B, 0, sizeof(B));
    std::vector<float4> vtab((size_t)nv);
    for (int v = 0; v < nv; v++) {
        // the sampler's per-slice arithmetic (sampleDistanceFieldT), same IEEE fp32 operations
        const float vf = (float)v;
        const int col = v / 3;
        const float rowIndex = std::floor(vf * g->invSliceCountXTimesOneThird);
        if (!(rowIndex >= 0.0f) || !(rowIndex < 65536.0f)) return ILB_OK;
        const int row = (int)rowIndex;
        const float cu = (float)col * g->sliceSizeX, rv = rowIndex * g->sliceSizeY;
        // extreme texel indices over cx in [0, ex], cy in [0, ey] (every step is monotonic in cx / cy)
        auto x0of = [&](float cx) { return std::floor((cu + cx * g->texelSizeX) * g->twf - 0.5f); };
        auto y0of = [&](float cy) { return std::floor((rv + cy * g->texelSizeY) * g->thf - 0.5f); };
        const float xa = x0of(0.0f), xb = x0of(g->ex), ya = y0of(0.0f), yb = y0of(g->ey);
        if (!(std::fabs(xa) < 1.0e6f) || !(std::fabs(xb) < 1.0e6f) || !(std::fabs(ya) < 1.0e6f) || !(std::fabs(yb) < 1.0e6f)) return ILB_OK;
        const int lx0 = (int)std::fmin(xa, xb) - col * sw, lx1 = (int)std::fmax(xa, xb) - col * sw;
        const int ly0 = (int)std::fmin(ya, yb) - row * sh, ly1 = (int)std::fmax(ya, yb) - row * sh;
        if (lx0 < -HALO || lx1 > sw + HALO - 1 || ly0 < -HALO || ly1 + 1 > sh + HALO - 1) return ILB_OK;
        B.col[v] = col; B.row[v] = row;
        const int base = v * pw * ph + (HALO - row * sh) * pw + (HALO - col * sw);
        // .w: the same base for indices that still carry the bias of floorBiased() in both coordinates (ilb_device.cuh), modulo 2^32
        const unsigned biased = (unsigned)base - (unsigned)ILB_FLOOR_BIAS * (unsigned)(pw + 1);
        float basef, biasedf;
        memcpy(&basef, &base, sizeof(float));
        memcpy(&biasedf, &biased, sizeof(float));
        vtab[(size_t)v] = make_float4(cu, rv, basef, biasedf);
    }

    if (stale && stale->nv == nv && stale->sw == sw && stale->sh == sh) {
        B.tex = df->tex; B.planes = stale->planes;
        B.tw = df->tw; B.th = df->th; B.sw = sw; B.sh = sh; B.pw = pw; B.ph = ph; B.nv = nv;
        launchPlaneBuild(ctx, B);
        ILB_CUDA(ctx, cudaGetLastError());
        stale->version = df->version;
        g->planes = stale->planes; g->vtab = stale->vtab; g->vtabBiased = biasedTable(stale->vtab); g->pitch = stale->pitch;
        return ILB_OK;
    }
    ilb_df_planes P;
    P.version = df->version; P.nv = nv; P.sw = sw; P.sh = sh;
    const size_t bytes = sizeof(float4) * (size_t)nv * pw * ph;
    if (cudaMalloc(&P.planes, bytes) != cudaSuccess) {
        cudaGetLastError();  // not enough HBM for the derived copy: keep sampling the atlas
        return ILB_OK;
    }
    if (cudaMalloc(&P.vtab, sizeof(float4) * (size_t)nv) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(P.planes);
        return ILB_OK;
    }
    cudaError_t e = cudaMemcpyAsync(P.vtab, vtab.data(), sizeof(float4) * (size_t)nv, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        B.tex = df->tex; B.planes = P.planes;
        B.tw = df->tw; B.th = df->th; B.sw = sw; B.sh = sh; B.pw = pw; B.ph = ph; B.nv = nv;
        launchPlaneBuild(ctx, B);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // vtab is a host vector
    if (e != cudaSuccess) {  // nothing was attached yet: release the derived copy before reporting
        cudaFree(P.planes);
        cudaFree(P.vtab);
        return ilb_cuda_fail(ctx, e, "build the expanded distance-field planes");
    }
    P.pitch = pw;
    const float key[10] = {g->sliceSizeX, g->sliceSizeY, g->texelSizeX, g->texelSizeY, g->ex, g->ey, g->maxValidZ, g->zToSlice,
                           g->invSliceCountXTimesOneThird, g->sliceCount};
    memcpy(P.key, key, sizeof(key));
    if (df->planes.size() >= 4) {  // at most a lighting set and a particle set are live in practice
        cudaFree(df->planes.front().planes);
        cudaFree(df->planes.front().vtab);
        df->planes.erase(df->planes.begin());
    }
    df->planes.push_back(P);
    g->planes = P.planes; g->vtab = P.vtab; g->vtabBiased = biasedTable(P.vtab); g->pitch = P.pitch;
    return ILB_OK;
}
