// Kernel (1)/(1'): per-Gaussian preprocess for the orthographic TSW camera (SURVEY.md Appendix A.1;
// call sites /root/reference/ortho_gaussian_renderer/renderer.py:90-98 and preprocess.py:99-104).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every product and sum below rounds
// separately, in the written order, so radii / tile rectangles / depth keys — the integers the
// binning stage is built from — are reproducible bit-for-bit by a scalar fp32 CPU restatement.
// The kernel is a pure stream (56 B in, <= 60 B out per Gaussian): HBM-bound, FMA fusion buys nothing.
#include "common.cuh"

namespace gsvc {

__device__ __forceinline__ float ldV(const DevSettings& s, int v, int r, int c)
{
    return __ldg(s.vt.V[v] + r * s.vt.vs_r[v] + c * s.vt.vs_c[v]);
}

// /root/reference/utils/sh_utils.py:26-43
__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

// Quaternion (r,x,y,z) -> rotation, convention of /root/reference/utils/general_utils.py:98-119,
// used as given (callers normalise: guassian.py:287).
__device__ __forceinline__ void quat_to_rot(const float4 q, float R[9])
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z);
    R[1] = 2.f * (x * y - r * z);
    R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);
    R[4] = 1.f - 2.f * (x * x + z * z);
    R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);
    R[7] = 2.f * (y * z + r * x);
    R[8] = 1.f - 2.f * (x * x + y * y);
}

__device__ __forceinline__ void cov3d_from_scale_rot(const float* s3, float mod, const float4 q, float cov[6])
{
    float R[9], M[9];
    quat_to_rot(q, R);
    const float sx = mod * s3[0], sy = mod * s3[1], sz = mod * s3[2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        M[3 * i + 0] = R[3 * i + 0] * sx;
        M[3 * i + 1] = R[3 * i + 1] * sy;
        M[3 * i + 2] = R[3 * i + 2] * sz;
    }
    cov[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    cov[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    cov[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    cov[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    cov[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    cov[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

// cov2D = scale^2 (W Sigma W^T)[0:2,0:2] + 0.3 I  — the orthographic Jacobian is scale*[I2|0] (no perspective term)
__device__ __forceinline__ void cov2d_ortho(const float cov[6], const float w0[3], const float w1[3], float scale,
                                            float& a, float& b, float& c)
{
    const float u00 = cov[0] * w0[0] + cov[1] * w0[1] + cov[2] * w0[2];
    const float u01 = cov[1] * w0[0] + cov[3] * w0[1] + cov[4] * w0[2];
    const float u02 = cov[2] * w0[0] + cov[4] * w0[1] + cov[5] * w0[2];
    const float u10 = cov[0] * w1[0] + cov[1] * w1[1] + cov[2] * w1[2];
    const float u11 = cov[1] * w1[0] + cov[3] * w1[1] + cov[4] * w1[2];
    const float u12 = cov[2] * w1[0] + cov[4] * w1[1] + cov[5] * w1[2];
    const float s2 = scale * scale;
    a = s2 * (w0[0] * u00 + w0[1] * u01 + w0[2] * u02) + LOWPASS;
    b = s2 * (w0[0] * u10 + w0[1] * u11 + w0[2] * u12);
    c = s2 * (w1[0] * u10 + w1[1] * u11 + w1[2] * u12) + LOWPASS;
}

__device__ __forceinline__ void sh_to_rgb(int deg, const float* sh, const float p[3], const float campos[3],
                                          float rgb[3], uint8_t clamped[3])
{
    const float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx * inv, y = dy * inv, z = dz * inv;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float r = SH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
                        SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
                }
            }
        }
        r += 0.5f;
        clamped[c] = (r < 0.0f);
        rgb[c] = r < 0.0f ? 0.0f : r;
    }
}

// Stable stream compaction of the visible anchors, fused into the filter kernel (what
// `anchor[radii_pure > 0]` — a nonzero + a host synchronisation — does after prefilter_voxel, preprocess.py:99-108,
// guassian.py:147-153).  Single pass, decoupled look-back: CTAs take a ticket (so logical order = scheduling
// order), publish their visible count, and resolve their exclusive prefix from their predecessors' published
// aggregates / inclusive prefixes.  state word: bits 63:62 = 0 invalid, 1 aggregate, 2 inclusive prefix.
struct CompactOut {
    int32_t* indices;                 // [P] ascending indices of the visible anchors (first `count` valid)
    unsigned long long* state;        // [n_ctas + 1]: [0] = ticket counter, [1 + b] = look-back word of logical CTA b
    unsigned long long* count_dev;    // device copy of the count (may be NULL)
    unsigned long long* host_slot;    // pinned host word: ticket << 40 | count (may be NULL)
    unsigned int ticket;
};
constexpr unsigned long long CPT_AGG = 1ull << 62, CPT_PREFIX = 2ull << 62, CPT_MASK = (1ull << 62) - 1;

__device__ __forceinline__ void compact_visible(const CompactOut& co, int cta, int n_ctas, int g, bool vis)
{
    __shared__ unsigned int s_wsum[8];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned int bal = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) s_wsum[wid] = __popc(bal);
    __syncthreads();
    unsigned int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const unsigned int x = s_wsum[w];
        if (w < wid) before += x;
        total += x;
    }
    if (wid == 0) {
        volatile unsigned long long* st = co.state + 1;
        if (lane == 0) st[cta] = (cta == 0 ? CPT_PREFIX : CPT_AGG) | total;
        unsigned long long excl = 0ull;
        int look = cta - 1;
        while (look >= 0) {                       // warp-wide look-back, 32 predecessors per round
            const int idx = look - lane;
            unsigned long long w = CPT_PREFIX;    // lanes before the first CTA: a zero inclusive prefix
            if (idx >= 0) { do { w = st[idx]; } while ((w >> 62) == 0ull); }
            const unsigned int has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
            const int stop = has_prefix ? __ffs(has_prefix) - 1 : 32;   // closest predecessor with an inclusive prefix
            unsigned long long v = lane <= stop ? (w & CPT_MASK) : 0ull;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            excl += v;
            if (has_prefix) break;
            look -= 32;
        }
        if (lane == 0) {
            if (cta > 0) st[cta] = CPT_PREFIX | (excl + total);
            s_excl = excl;
            if (cta == n_ctas - 1) {
                const unsigned long long cnt = excl + total;
                if (co.count_dev) *co.count_dev = cnt;
                if (co.host_slot) *co.host_slot = ((unsigned long long)co.ticket << 40) | cnt;
            }
        }
    }
    __syncthreads();
    if (vis) co.indices[s_excl + before + __popc(bal & ((1u << lane) - 1u))] = g;
}

// MODE 0: full preprocess + per-tile instance counting.  MODE 1: visible_filter (radii only).
// MODE 2: visible_filter + stable compaction of the visible indices.
template <int MODE>
__global__ void __launch_bounds__(256) preprocess_kernel(DevSettings s, PreInputs in, int32_t* __restrict__ radii,
                                                         GeomView geo, unsigned int* __restrict__ tile_count,
                                                         float4* __restrict__ acc_to_zero, CompactOut co)
{
    constexpr bool FILTER = MODE != 0;
    pdl_prologue();
    // Lanes past the end stay in the warp (they redo the last Gaussian with all stores masked) so that the
    // warp-wide tile walk at the end (and the block-wide compaction) runs converged.
    __shared__ int s_cta;
    int cta = blockIdx.x;
    if (MODE == 2) {                              // logical CTA index = ticket: look-back never waits on a later CTA
        if (threadIdx.x == 0) s_cta = (int)atomicAdd(co.state, 1ull);
        __syncthreads();
        cta = s_cta;
    }
    const int g_raw = cta * blockDim.x + threadIdx.x;
    const bool valid = g_raw < in.P;
    if (MODE == 1 && !valid) return;
    const int g = valid ? g_raw : in.P - 1;
    // view of the batch (grid.y): per-Gaussian state of view v lives at virtual index v*P + g, its tiles at v*Tv + t
    const int v = FILTER ? 0 : (int)blockIdx.y;
    const size_t gv = (size_t)v * in.P + g;
    if (!FILTER && acc_to_zero && valid) {
        // the blend backward accumulates into these 9 sums with atomics: zero them here (a store the
        // stream kernel hides) instead of a separate memset launch in the backward
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        acc_to_zero[3 * gv] = z; acc_to_zero[3 * gv + 1] = z; acc_to_zero[3 * gv + 2] = z;
    }

    const float p[3] = {__ldg(in.means3D + 3 * g), __ldg(in.means3D + 3 * g + 1), __ldg(in.means3D + 3 * g + 2)};
    // scales / quaternion are fetched together with the position (one memory round trip per Gaussian, at the
    // price of 28 wasted bytes for the third of the Gaussians the slab test culls)
    float sc[3] = {0.f, 0.f, 0.f};
    float4 q_pre = make_float4(1.f, 0.f, 0.f, 0.f);
    if (!in.cov3D_precomp) {
        sc[0] = __ldg(in.scales + 3 * g); sc[1] = __ldg(in.scales + 3 * g + 1); sc[2] = __ldg(in.scales + 3 * g + 2);
        q_pre = __ldg(reinterpret_cast<const float4*>(in.rotations) + g);
    }
    const float w0[3] = {ldV(s, v, 0, 0), ldV(s, v, 0, 1), ldV(s, v, 0, 2)};
    const float w1[3] = {ldV(s, v, 1, 0), ldV(s, v, 1, 1), ldV(s, v, 1, 2)};
    const float vx = w0[0] * p[0] + w0[1] * p[1] + w0[2] * p[2] + ldV(s, v, 0, 3);
    const float vy = w1[0] * p[0] + w1[1] * p[1] + w1[2] * p[2] + ldV(s, v, 1, 3);
    const float vz = ldV(s, v, 2, 0) * p[0] + ldV(s, v, 2, 1) * p[1] + ldV(s, v, 2, 2) * p[2] + ldV(s, v, 2, 3);

    int radius = 0;
    int rminx = 0, rminy = 0, rmaxx = 0, rmaxy = 0;
    float a = 0.f, b = 0.f, c = 0.f, det = 0.f, px = 0.f, py = 0.f;
    // U6: TSW slab — keep |z_view| <= threshold (preprocess.py:109-116)
    bool ok = valid && !(fabsf(vz) > s.threshold);
    if (ok) {
        float cov[6];
        if (in.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov[k] = __ldg(in.cov3D_precomp + 6 * (size_t)g + k);
        } else {
            cov3d_from_scale_rot(sc, s.scale_modifier, q_pre, cov);
        }
        cov2d_ortho(cov, w0, w1, s.scale, a, b, c);
        det = a * c - b * b;
        ok = (det != 0.0f);
    }
    if (ok) {
        const float mid = 0.5f * (a + c);
        const float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        const float rad_f = ceilf(3.f * sqrtf(lam));
        radius = (int)rad_f;
        px = (vx - s.x_min) * s.scale - 0.5f;  // U1
        py = (vy - s.y_min) * s.scale - 0.5f;
        const float fgx = (float)s.gx, fgy = (float)s.gy, ft = (float)TILE;
        rminx = (int)fminf(fgx, fmaxf(0.f, truncf((px - rad_f) / ft)));
        rminy = (int)fminf(fgy, fmaxf(0.f, truncf((py - rad_f) / ft)));
        rmaxx = (int)fminf(fgx, fmaxf(0.f, truncf((px + rad_f + (float)(TILE - 1)) / ft)));
        rmaxy = (int)fminf(fgy, fmaxf(0.f, truncf((py + rad_f + (float)(TILE - 1)) / ft)));
        ok = (rmaxx - rminx) * (rmaxy - rminy) > 0;
    }
    if (valid && radii) radii[gv] = ok ? radius : 0;
    if (MODE == 2) compact_visible(co, cta, (int)gridDim.x, g, ok);
    if (FILTER) return;

    if (ok) {
        const float det_inv = 1.f / det;
        const float cA = c * det_inv, cB = -b * det_inv, cC = a * det_inv;
        const float op = __ldg(in.opacities + g);
        float rgb[3];
        if (in.colors_precomp) {
            rgb[0] = __ldg(in.colors_precomp + 3 * g);
            rgb[1] = __ldg(in.colors_precomp + 3 * g + 1);
            rgb[2] = __ldg(in.colors_precomp + 3 * g + 2);
        } else {
            uint8_t cl[3];
            const float campos[3] = {s.vt.campos[v][0], s.vt.campos[v][1], s.vt.campos[v][2]};
            sh_to_rgb(s.sh_degree, in.shs + (size_t)g * s.sh_M * 3, p, campos, rgb, cl);
            geo.clamped[3 * gv] = cl[0];
            geo.clamped[3 * gv + 1] = cl[1];
            geo.clamped[3 * gv + 2] = cl[2];
        }
        geo.feat0[gv] = make_float4(px, py, 0.f, 0.f);   // .zw: scratch of the blend kernels' staging (B/A, B/C)
        geo.feat1[gv] = make_float4(cA, cB, cC, op);
        geo.feat2[gv] = make_float4(rgb[0], rgb[1], rgb[2], vz);
        geo.rect[gv] = make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx,
                                    (unsigned short)rmaxy);
    } else if (valid) {
        geo.rect[gv] = make_ushort4(0, 0, 0, 0);
    }
    // per-tile instance histogram (the counting sort's digit): the warp walks the flattened list of all its
    // (Gaussian, tile) pairs, 32 at a time, so every RED instruction is full width whatever the rectangle sizes
    const int rw = rmaxx - rminx;
    const WarpTiles wt = warp_tiles_begin(ok ? rw * (rmaxy - rminy) : 0, rminx, rminy, rw > 0 ? rw : 1);
    for (int base = 0; base < wt.total; base += 32) {
        int owner;
        const int t = warp_tiles_get(wt, base + (threadIdx.x & 31), s.gx, owner);
        if (t >= 0) atomicAdd(tile_count + (size_t)v * (s.gx * s.gy) + t, 1u);
    }
    // Measured and rejected (profiles/r1_notes.md): (a) taking each instance's rank inside its tile from this
    // atomic's return value and writing (composite, tile, rank) records so that the scatter needs no second atomic
    // pass; (b) running the tile scan in the last CTA of this kernel (threadfence + counter).  Both made the
    // 2-view step slower (+15 us): this kernel is instruction-bound, the fire-and-forget RED is free for it while
    // a returning ATOM or a serial single-CTA scan tail is not.

}

cudaError_t launch_visible_filter(const DevSettings& s, const PreInputs& in, int32_t* radii, cudaStream_t st)
{
    if (in.P <= 0) return cudaSuccess;
    GeomView none{};
    count_launch();
    return launch_pdl(preprocess_kernel<1>, dim3((in.P + 255) / 256), dim3(256), st, s, in, radii, none,
                      (unsigned int*)nullptr, (float4*)nullptr, CompactOut{});
}

cudaError_t launch_visible_filter_compact(const DevSettings& s, const PreInputs& in, int32_t* radii, int32_t* indices,
                                          void* scratch, unsigned long long* host_slot, unsigned int ticket,
                                          cudaStream_t st)
{
    const int n_ctas = (in.P + 255) / 256;
    CompactOut co;
    co.indices = indices;
    co.state = static_cast<unsigned long long*>(scratch) + 1;      // [0] of scratch = device copy of the count
    co.count_dev = static_cast<unsigned long long*>(scratch);
    co.host_slot = host_slot;
    co.ticket = ticket;
    cudaError_t e = cudaMemsetAsync(scratch, 0, compact_scratch_bytes(in.P), st);
    if (e != cudaSuccess) return e;
    if (in.P <= 0) {
        if (host_slot) *host_slot = (unsigned long long)ticket << 40;   // pinned HOST memory: count 0, no launch
        return cudaSuccess;
    }
    GeomView none{};
    count_launch();
    return launch_pdl(preprocess_kernel<2>, dim3(n_ctas), dim3(256), st, s, in, radii, none, (unsigned int*)nullptr,
                      (float4*)nullptr, co);
}

cudaError_t launch_preprocess(const DevSettings& s, const PreInputs& in, int32_t* radii, GeomView g, ImageView im,
                              float4* acc_to_zero, cudaStream_t st)
{
    // one memset clears the per-tile counters and the scan's per-CTA partials that sit right behind them
    const size_t T = (size_t)s.gx * s.gy * s.n_views;
    const size_t nbytes = reinterpret_cast<char*>(im.scan_partials + (T / 1024 + 1)) - reinterpret_cast<char*>(im.tile_count);
    cudaError_t e = cudaMemsetAsync(im.tile_count, 0, nbytes, st);
    if (e != cudaSuccess) return e;
    if (in.P <= 0) return cudaSuccess;
    count_launch();
    return launch_pdl(preprocess_kernel<0>, dim3((in.P + 255) / 256, s.n_views), dim3(256), st, s, in, radii, g,
                      im.tile_count, acc_to_zero, CompactOut{});
}

}  // namespace gsvc
