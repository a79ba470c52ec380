// Kernels (3) and (4): front-to-back alpha blend forward and its backward
// (the reference reaches them through
//  /root/reference/ortho_gaussian_renderer/renderer.py:90-98 and loss.backward(), pipeline/train.py:462).
//
// One CTA (128 threads) per 16x16 tile; each of its 4 warps owns an 8x8-pixel block and every lane
// two pixels of it (rows r and r+4 of one column).  Gaussians of the tile's depth-sorted list are
// gathered with float4 loads into shared memory in batches of 256.  Both kernels are
// instruction-issue bound (ncu: ~85 % issue-active, < 5 % DRAM), so the design goal is to execute
// fewer and cheaper pixel-Gaussian evaluations, not to move fewer bytes:
//   * exact sub-tile culling — per chunk of 32 staged Gaussians every lane tests ONE Gaussian: can
//     its alpha >= 1/255 ellipse reach the warp's 8x8 block (closed-form maximum of the concave
//     exponent over the rectangle)?  A ballot gives the warp its private work list.  Only pairs the
//     reference would `continue` past (alpha < 1/255, no state change) are skipped, so results are
//     unchanged;
//   * two pixels per lane share the Gaussian fetch, the x terms of the exponent, the loop control and
//     — in the backward — the warp reduction, which is the single most expensive step;
//   * the conic is pre-scaled by -log2(e)/2 while staging, so a pair costs ~5 FP32 ops + one
//     ex2.approx (MUFU) up to the alpha test; staged records are 48 B interleaved so one address
//     serves all three LDS.128;
//   * warp-ballot early termination once all 64 pixels of a block are saturated (T < 1e-4 stop rule);
//   * backward: branch-free per-pair arithmetic on raw moments (the per-Gaussian kernel finishes
//     dL/dpix and dL/dconic), 9 sums reduced across the warp with a transpose-butterfly (14 shuffles
//     instead of 45), and 9 lanes issue the 9 atomics of one Gaussian in a single instruction.
#include <cstdlib>

#include "common.cuh"

namespace gsvc {

constexpr int BLEND_THREADS = 128;  // 4 warps x (8x8 pixels), 2 pixels per lane
constexpr int BLEND_WARPS = BLEND_THREADS / 32;
constexpr int BATCH = 256;          // Gaussians staged per round (2 per thread)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) ---------------------------------------------
// Each lane blends two pixels (A = low half, B = high half) with identical instruction streams, and both
// blend kernels are bound by instruction issue, not by the FP32 pipes (ncu: ~85 % issue-active, fma pipe
// ~35 %).  One f32x2 instruction does the work of two scalar ones in a single issue slot (measured on
// B200: scripts/ubench/ffma2.cu); ptxas folds broadcast scalars, immediates and negations into its operands.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 bc2(float s) { return mk2(s, s); }
__device__ __forceinline__ float lo(f2 a)
{
    float x;
    [[maybe_unused]] float y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return x;
}
__device__ __forceinline__ float hi(f2 a)
{
    [[maybe_unused]] float x;
    float y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return y;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
// c += a * b in place (ties the accumulator to one register pair across loop iterations)
__device__ __forceinline__ void fma2_acc(f2& c, f2 a, f2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c.v) : "l"(a.v), "l"(b.v));
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-lo(a), -hi(a)); }

// MINUS the log2 of the falloff of one staged Gaussian at this lane's two pixels, as a sum of squares:
//   -q = (l11 (dx + rho dy))^2 + (l22 dy)^2,   L = (l11 0; l11 rho, l22) = Cholesky factor of the conic scaled by log2(e)/2
// (dx is shared: same column).  Unlike A dx^2 + B dx dy + C dy^2 nothing cancels — for a needle-like Gaussian far
// from its centre those three terms are ~1e2 each and sum to ~1, which costs ~1e-5 of absolute accuracy in fp32 —
// and -q >= 0 holds by construction, so the reference's `power > 0 -> continue` (only ever taken through rounding)
// needs no test.
// rho = B/A comes as rho_hi + rho_lo (see preprocess.cu): dx + rho dy cancels along an elongated Gaussian, so the
// shear is accumulated with two FMAs (exact products, rounding at the small result) before the scale by l11.
__device__ __forceinline__ f2 neg_falloff_log2(float rho_hi, const float4 f1, float dx, f2 dy2, f2& u2, f2& v2)
{
    const f2 t2 = fma2(bc2(f1.y), dy2, fma2(bc2(rho_hi), dy2, bc2(dx)));
    u2 = mul2(bc2(f1.x), t2);
    v2 = mul2(bc2(f1.z), dy2);
    return fma2(u2, u2, mul2(v2, v2));
}
__device__ __forceinline__ f2 neg_falloff_log2(float rho_hi, const float4 f1, float dx, f2 dy2)
{
    f2 u2, v2;
    return neg_falloff_log2(rho_hi, f1, dx, dy2, u2, v2);
}

struct BlockGeom {
    int px, py;                    // this lane's first pixel; the second one is (px, py + 4)
    float xmin, xmax, ymin, ymax;  // pixel-centre bounds of the warp's 8x8 block
};

__device__ __forceinline__ BlockGeom block_geom(int tile, int gx, int tid)
{
    const int tx = tile % gx, ty = tile / gx;
    const int warp = tid >> 5, lane = tid & 31;
    const int bx0 = tx * TILE + (warp & 1) * 8, by0 = ty * TILE + (warp >> 1) * 8;
    BlockGeom b;
    b.px = bx0 + (lane & 7);
    b.py = by0 + (lane >> 3);
    b.xmin = (float)bx0; b.xmax = (float)(bx0 + 7);
    b.ymin = (float)by0; b.ymax = (float)(by0 + 7);
    return b;
}

// Staged record k = s_feat[3k .. 3k+2]:
//   [0] = (pix.x, pix.y, rho_hi, B/C)        rho = B/A = rho_hi + rho_lo
//   [1] = (l11, rho_lo, l22, opacity): Cholesky factor (l11 0; l11 rho, l22) of (A, B; B, C) * log2(e)/2, so that
//         alpha = opacity * 2^q(d),  -q(d) = (l11 (dx + rho dy))^2 + (l22 dy)^2
//   [2] = (r, g, b, Gaussian id as bits)
constexpr float CULL_MARGIN = 1e-3f;  // in log2 units (7e-4 relative in alpha) >> fp32 rounding of q

__device__ __forceinline__ void stage(float4* s_feat, int slot, const GeomView& geo, unsigned int id)
{
    // everything was prepared per Gaussian by the preprocess kernel: three 16-byte gathers, no arithmetic
    const float4 f0 = __ldg(geo.feat0 + id);
    const float4 f1 = __ldg(geo.feat3 + id);
    float4 f2 = __ldg(geo.feat2 + id);
    f2.w = __uint_as_float(id);   // rides along with the colour: the backward's atomics need it per hit
    s_feat[3 * slot] = f0;
    s_feat[3 * slot + 1] = f1;
    s_feat[3 * slot + 2] = f2;
}

// Exact sub-tile culling: can the staged Gaussian reach alpha >= 1/255 anywhere in the block
// [xmin,xmax] x [ymin,ymax]?  q is concave with its maximum (0) at the centre, so its maximum over
// the rectangle is attained at the centre if that lies inside, else on an edge facing the centre:
// on the vertical line x = clamp(cx) the maximiser is y = cy - (B/C)(x - cx), on the horizontal
// line y = clamp(cy) it is x = cx - (B/A)(y - cy), each clamped to the edge.  Both candidates are
// points of the rectangle, so max(q1, q2) never over-estimates, and it equals the true maximum.
__device__ __forceinline__ bool block_hit(const BlockGeom& b, const float4* e)
{
    const float4 f0 = e[0];
    const float4 f1 = e[1];
    const float thr = -__log2f(255.0f * f1.w) - CULL_MARGIN;   // alpha >= 1/255  <=>  q >= thr (+margin)
    const float ex = fminf(fmaxf(f0.x, b.xmin), b.xmax), ey = fminf(fmaxf(f0.y, b.ymin), b.ymax);
    const float dxe = ex - f0.x, dye = ey - f0.y;
    const float dy1 = fminf(fmaxf(fmaf(-f0.w, dxe, f0.y), b.ymin), b.ymax) - f0.y;
    const float dx2 = fminf(fmaxf(fmaf(-f0.z, dye, f0.x), b.xmin), b.xmax) - f0.x;
    const float u1 = f1.x * fmaf(f0.z, dy1, dxe), v1 = f1.z * dy1;     // (rho_lo is far below the culling margin)
    const float u2 = f1.x * fmaf(f0.z, dye, dx2), v2 = f1.z * dye;
    const float nq1 = fmaf(u1, u1, v1 * v1), nq2 = fmaf(u2, u2, v2 * v2);   // -q at the two candidates
    return -fminf(nq1, nq2) >= thr;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLEND_THREADS)
render_forward_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, unsigned long long cap,
                      float* __restrict__ out_color)
{
    __shared__ float4 s_feat[3 * BATCH];

    pdl_prologue();
    // virtual tile blockIdx.x = view * Tv + tile (a batch of views is one launch; the drop-in call has one view)
    const int Tv = s.gx * s.gy;
    const int view = (int)blockIdx.x / Tv;
    const int tile = (int)blockIdx.x - view * Tv;
    const int tid = threadIdx.x, lane = tid & 31;
    const BlockGeom bg = block_geom(tile, s.gx, tid);
    const bool inA = bg.px < s.W && bg.py < s.H, inB = bg.px < s.W && bg.py + 4 < s.H;
    const float pxf = (float)bg.px, pyf = (float)bg.py;

    const uint2 rg = im.ranges[blockIdx.x];
    if ((unsigned long long)rg.y > cap) return;  // capacity overflow: host re-runs with a larger buffer
    const int n = (int)(rg.y - rg.x);

    // pixel A = (px, py) in the low halves, pixel B = (px, py + 4) in the high halves
    float TA = 1.f, A0 = 0.f, A1 = 0.f, A2 = 0.f;
    float TB = 1.f, B0 = 0.f, B1 = 0.f, B2 = 0.f;
    const f2 npy2 = mk2(-pyf, -(pyf + 4.f));
    unsigned int lastA = 0, lastB = 0;
    // A pixel that is finished (outside the image, or stopped by the T < 1e-4 rule) is represented by an
    // unreachable alpha floor, so "still live" costs no instruction in the pair loop: the reference's
    // alpha >= 1/255 test is made against this per-pixel register.
    constexpr float DONE = 2.f;   // alpha <= 0.99 < DONE
    float minA = inA ? ALPHA_MIN : DONE, minB = inB ? ALPHA_MIN : DONE;

    for (int base = 0; base < n; base += BATCH) {
        // block-wide vote doubles as the barrier that protects the staging buffer
        if (__syncthreads_count(minA > 1.f && minB > 1.f) == BLEND_THREADS) break;
#pragma unroll
        for (int r = 0; r < BATCH / BLEND_THREADS; r++) {
            const int slot = tid + r * BLEND_THREADS;
            if (base + slot < n) stage(s_feat, slot, geo, bin.point_list[rg.x + base + slot]);
        }
        __syncthreads();
        const int cnt = min(BATCH, n - base);
        for (int c = 0; c < cnt; c += 32) {
            // warp-ballot early termination: all 64 pixels of the block saturated
            if (__ballot_sync(FULL, !(minA > 1.f && minB > 1.f)) == 0u) break;
            bool hit = false;
            if (c + lane < cnt) hit = block_hit(bg, s_feat + 3 * (c + lane));
            unsigned int m = __ballot_sync(FULL, hit);
            const float4* chunk = s_feat + 3 * c;
            const unsigned int pos1 = (unsigned int)(base + c + 1);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const float4* e = chunk + 3 * k;
                const float4 f0 = e[0];
                const float4 f1 = e[1];
                const float4 f2v = e[2];
                const f2 nq2 = neg_falloff_log2(f0.z, f1, f0.x - pxf, add2(bc2(f0.y), npy2));
                const f2 a2 = mul2(bc2(f1.w), mk2(ex2_approx(-lo(nq2)), ex2_approx(-hi(nq2))));
                const float aA = fminf(ALPHA_MAX, lo(a2)), aB = fminf(ALPHA_MAX, hi(a2));
                // the reference `continue`s on alpha < 1/255 (power > 0 cannot happen here; a finished pixel's
                // floor is unreachable)
                const bool okA = !(aA < minA);
                const bool okB = !(aB < minB);
                const unsigned int pos = pos1 + (unsigned int)k;
                const f2 ac = mk2(aA, aB), T2 = mk2(TA, TB);
                const f2 tT2 = mul2(T2, sub2(bc2(1.f), ac));
                const f2 w2 = mul2(ac, T2);
                const float tTA = lo(tT2), tTB = hi(tT2);
                // T' < 1e-4: the pixel stops and this Gaussian is NOT added
                if (okA && tTA < T_STOP) minA = DONE;
                if (okB && tTB < T_STOP) minB = DONE;
                const bool addA = okA && !(tTA < T_STOP), addB = okB && !(tTB < T_STOP);
                const f2 wz = mk2(addA ? lo(w2) : 0.f, addB ? hi(w2) : 0.f);   // adding 0 leaves the colour bit-identical
                { const f2 c = fma2(bc2(f2v.x), wz, mk2(A0, B0)); A0 = lo(c); B0 = hi(c); }
                { const f2 c = fma2(bc2(f2v.y), wz, mk2(A1, B1)); A1 = lo(c); B1 = hi(c); }
                { const f2 c = fma2(bc2(f2v.z), wz, mk2(A2, B2)); A2 = lo(c); B2 = hi(c); }
                if (addA) { TA = tTA; lastA = pos; }
                if (addB) { TB = tTB; lastB = pos; }
            }
        }
    }
    const size_t N = (size_t)s.W * s.H;
    const float bg0 = __ldg(s.bg + 0), bg1 = __ldg(s.bg + 1), bg2 = __ldg(s.bg + 2);
    // the view's own per-pixel state (unflipped), and its target: image out_image, optionally mirrored in x,
    // scaled by weight; views that share an image (toast halves) accumulate with atomics onto a zeroed image
    float* const fT = im.final_T + (size_t)view * N;
    unsigned int* const nc = im.n_contrib + (size_t)view * N;
    float* const out = out_color + (size_t)s.vt.out_image[view] * 3 * N;
    const float wgt = s.vt.weight[view];
    const int ox = s.vt.flip_x[view] ? s.W - 1 - bg.px : bg.px;
    if (inA) {
        const size_t pix = (size_t)bg.py * s.W + bg.px, opix = (size_t)bg.py * s.W + ox;
        const float c0 = wgt * (A0 + TA * bg0), c1 = wgt * (A1 + TA * bg1), c2 = wgt * (A2 + TA * bg2);
        if (s.accumulate) {
            atomicAdd(out + opix, c0); atomicAdd(out + N + opix, c1); atomicAdd(out + 2 * N + opix, c2);
        } else {
            out[opix] = c0; out[N + opix] = c1; out[2 * N + opix] = c2;
        }
        fT[pix] = TA;
        nc[pix] = lastA;
    }
    if (inB) {
        const size_t pix = (size_t)(bg.py + 4) * s.W + bg.px, opix = (size_t)(bg.py + 4) * s.W + ox;
        const float c0 = wgt * (B0 + TB * bg0), c1 = wgt * (B1 + TB * bg1), c2 = wgt * (B2 + TB * bg2);
        if (s.accumulate) {
            atomicAdd(out + opix, c0); atomicAdd(out + N + opix, c1); atomicAdd(out + 2 * N + opix, c2);
        } else {
            out[opix] = c0; out[N + opix] = c1; out[2 * N + opix] = c2;
        }
        fT[pix] = TB;
        nc[pix] = lastB;
    }
}

cudaError_t launch_render_forward(const DevSettings& s, GeomView g, ImageView im, BinView b, long long cap,
                                  float* out_color, cudaStream_t st)
{
    const int T = s.gx * s.gy * s.n_views;
    if (T <= 0) return cudaSuccess;
    count_launch();
    return launch_pdl(render_forward_kernel, dim3(T), dim3(BLEND_THREADS), st, s, g, im, b, (unsigned long long)cap,
                      out_color);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Sum 8 values across the warp with 9 shuffles: after three exchange-and-halve steps every lane
// holds one partial, after two more butterfly steps lane L holds the full sum of value index
// (L >> 2) in v[0].
__device__ __forceinline__ float warp_reduce8(float v[8], int lane)
{
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float send = hi ? v[i] : v[i + 4];
        const float keep = hi ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = hi ? v[i] : v[i + 2];
        const float keep = hi ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    hi = lane & 4;
    {
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 2);
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    return v[0];
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// atomicAdd(a, v) on the lanes where `on` is non-zero, as ONE predicated instruction: written as a branch the compiler
// wraps the address arithmetic and the reduction in a BSSY / BRA / BSYNC region (4 extra issue slots per hit).
__device__ __forceinline__ void red_add_if(int on, float* a, float v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q red.global.add.f32 [%0], %1;\n\t}"
                 ::"l"(a), "f"(v), "r"(on) : "memory");
}

// Per-lane replay state of the backward: both pixels of the lane (A = low half, B = high half).
//   T   running transmittance (T_i before the current Gaussian)
//   Sg  g . (colour composited behind the current Gaussian, incl. T_final * bg):
//         Sg_i = sum_{j>i} (c_j . g) alpha_j T_j + T_final (bg . g)
//   g*  dL/dC of the pixel
// With dC/dalpha_i = c_i T_i - (sum_{j>i} c_j alpha_j T_j + T_final bg) / (1 - alpha_i) contracted with
// g = dL/dC up front, the behind-colour recursion is ONE scalar per pixel (Sg) instead of three channels.
struct PairBwd {
    float TA, TB, SgA, SgB;
    f2 g0, g1, g2;
    unsigned int lastA, lastB;   // n_contrib
};

__global__ void __launch_bounds__(BLEND_THREADS)
render_backward_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, unsigned long long cap,
                       const float* __restrict__ dL_dout, float* __restrict__ acc /* [P][12] */)
{
    __shared__ float4 s_feat[3 * BATCH];
    __shared__ unsigned int s_max[BLEND_WARPS];

    pdl_prologue();
    const int Tv = s.gx * s.gy;
    const int view = (int)blockIdx.x / Tv;          // virtual tile = view * Tv + tile
    const int tile = (int)blockIdx.x - view * Tv;
    const int tid = threadIdx.x, lane = tid & 31;
    const BlockGeom bg = block_geom(tile, s.gx, tid);
    const bool inA = bg.px < s.W && bg.py < s.H, inB = bg.px < s.W && bg.py + 4 < s.H;
    const float pxf = (float)bg.px, pyf = (float)bg.py;
    const size_t N = (size_t)s.W * s.H;
    const size_t pixA = (size_t)bg.py * s.W + bg.px, pixB = (size_t)(bg.py + 4) * s.W + bg.px;
    // dL/d(view image) = weight * dL/d(output image), read mirrored in x for a flipped view
    const int ox = s.vt.flip_x[view] ? s.W - 1 - bg.px : bg.px;
    const size_t opixA = (size_t)bg.py * s.W + ox, opixB = (size_t)(bg.py + 4) * s.W + ox;
    const float wgt = s.vt.weight[view];
    const float* const dL = dL_dout + (size_t)s.vt.out_image[view] * 3 * N;
    const float* const fT = im.final_T + (size_t)view * N;
    const unsigned int* const nc = im.n_contrib + (size_t)view * N;

    const uint2 rg = im.ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0) return;
    // Capacity overflow (only a CUDA-graph replay can get here with one: the eager caller re-runs the forward on a
    // larger buffer first): the forward skipped the overflowed tiles, so their final_T / n_contrib are stale and the
    // point list beyond `cap` does not exist.  The whole launch is void — the frame is invalid and its owner is told
    // so by capacity_ok() — hence every tile leaves before it reads a list entry.
    if ((unsigned long long)rg.y > cap || im.hdr->overflow != 0u) return;

    // The tile's start-up is a chain of dependent memory round trips: ranges -> n_contrib (-> deepest contributor) ->
    // point list -> per-Gaussian records -> shared memory.  Where the replay starts depends on n_contrib, but for most
    // tiles it starts at the end of the list (nothing saturated): the ids of the last BATCH entries are fetched
    // speculatively right away, in parallel with the pixel state, which takes one round trip out of the chain.
    unsigned int spec_id[BATCH / BLEND_THREADS];
#pragma unroll
    for (int r = 0; r < BATCH / BLEND_THREADS; r++) {
        const int kp = n - 1 - (tid + r * BLEND_THREADS);
        spec_id[r] = kp >= 0 ? __ldg(bin.point_list + rg.x + kp) : 0u;
    }

    const float bg0 = __ldg(s.bg), bg1 = __ldg(s.bg + 1), bg2 = __ldg(s.bg + 2);
    PairBwd S{};
    {
        float gA0 = 0.f, gA1 = 0.f, gA2 = 0.f, gB0 = 0.f, gB1 = 0.f, gB2 = 0.f;
        if (inA) {
            S.TA = fT[pixA]; S.lastA = nc[pixA];
            gA0 = wgt * dL[opixA]; gA1 = wgt * dL[N + opixA]; gA2 = wgt * dL[2 * N + opixA];
            S.SgA = S.TA * (bg0 * gA0 + bg1 * gA1 + bg2 * gA2);
        }
        if (inB) {
            S.TB = fT[pixB]; S.lastB = nc[pixB];
            gB0 = wgt * dL[opixB]; gB1 = wgt * dL[N + opixB]; gB2 = wgt * dL[2 * N + opixB];
            S.SgB = S.TB * (bg0 * gB0 + bg1 * gB1 + bg2 * gB2);
        }
        S.g0 = mk2(gA0, gB0); S.g1 = mk2(gA1, gB1); S.g2 = mk2(gA2, gB2);
    }
    const f2 npy2 = mk2(-pyf, -(pyf + 4.f));

    // nothing behind the deepest contributor of the tile matters: start the replay there
    const unsigned int wmax = __reduce_max_sync(FULL, max(S.lastA, S.lastB));
    if (lane == 0) s_max[tid >> 5] = wmax;
    __syncthreads();
    unsigned int bmax = 0;
#pragma unroll
    for (int w = 0; w < BLEND_WARPS; w++) bmax = max(bmax, s_max[w]);
    const int m_len = min((int)bmax, n);  // list entries [0, m_len) are replayed, back to front

    // lanes 0,4,..,28 own the 8 reduced sums (index lane/4), lane 1 the ninth: one atomic instruction
    const int red_lane = ((lane & 3) == 0 || lane == 1) ? 1 : 0;
    const int red_off = lane == 1 ? 8 : (lane >> 2);

    for (int base = 0; base < m_len; base += BATCH) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < BATCH / BLEND_THREADS; r++) {
            const int slot = tid + r * BLEND_THREADS;
            const int kpos = m_len - 1 - (base + slot);  // list position staged in this slot
            if (kpos >= 0) {
                const unsigned int id = (base == 0 && m_len == n) ? spec_id[r] : bin.point_list[rg.x + kpos];
                stage(s_feat, slot, geo, id);
            }
        }
        __syncthreads();
        const int cnt = min(BATCH, m_len - base);
        for (int c = 0; c < cnt; c += 32) {
            // staged slot j holds list position m_len-1-(base+j); only positions below the warp's deepest
            // contributor can matter
            const int pos0 = m_len - 1 - (base + c);  // list position of chunk slot 0
            bool hit = false;
            if (c + lane < cnt && (unsigned int)(pos0 - lane) < wmax) hit = block_hit(bg, s_feat + 3 * (c + lane));
            unsigned int m = __ballot_sync(FULL, hit);
            const float4* chunk = s_feat + 3 * c;
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const float4* e = chunk + 3 * k;
                const float4 f0 = e[0];
                const float4 f1 = e[1];
                const float4 f2v = e[2];
                const float dx = f0.x - pxf;
                const f2 dy2 = add2(bc2(f0.y), npy2);
                f2 u2, v2;                                                // (u, v) = L^T d: the whitened offset
                const f2 nq2 = neg_falloff_log2(f0.z, f1, dx, dy2, u2, v2);
                const f2 Gs2 = mk2(ex2_approx(-lo(nq2)), ex2_approx(-hi(nq2)));
                const f2 a2 = mul2(bc2(f1.w), Gs2);                       // opacity * Gs, before the 0.99 cap
                const unsigned int pos = (unsigned int)(pos0 - k);
                // min(0.99, a) < 1/255  <=>  a < 1/255: the floor is tested on the uncapped value
                const bool useA = (pos < S.lastA) && !(lo(a2) < ALPHA_MIN);
                const bool useB = (pos < S.lastB) && !(hi(a2) < ALPHA_MIN);
                // (no warp vote here: after the exact culling above 99.9 % of the evaluations that reach this point
                //  contribute to at least one pixel — ncu source counters — and a pair that does not adds zeros)

                // Branch-free: a pair that does not contribute is blended with alpha = 0, which leaves T, the
                // behind-colour sum and every gradient sum unchanged (bit-identically).
                // The masked UNCAPPED value am does double duty: capped it is the blended alpha, and times
                // dL/dalpha it is w = opacity * Gs * dL/dalpha (U4: straight through the cap), already zero for a
                // pair that does not contribute — no separate masking of dL/dalpha.
                const f2 am2 = mk2(useA ? lo(a2) : 0.f, useB ? hi(a2) : 0.f);
                const f2 ae2 = mk2(fminf(ALPHA_MAX, lo(am2)), fminf(ALPHA_MAX, hi(am2)));
                const f2 om2 = sub2(bc2(1.f), ae2);
                const f2 ra2 = mk2(rcp_approx(lo(om2)), rcp_approx(hi(om2)));
                const f2 T2 = mul2(mk2(S.TA, S.TB), ra2);                 // T_i = T_{i+1} / (1 - alpha_i)
                S.TA = lo(T2); S.TB = hi(T2);
                const f2 cg2 = fma2(bc2(f2v.x), S.g0, fma2(bc2(f2v.y), S.g1, mul2(bc2(f2v.z), S.g2)));
                const f2 Sg2 = mk2(S.SgA, S.SgB);
                const f2 dla2 = fma2(T2, cg2, neg2(mul2(ra2, Sg2)));       // dL/dalpha
                const f2 dchan2 = mul2(ae2, T2);
                { const f2 n = fma2(dchan2, cg2, Sg2); S.SgA = lo(n); S.SgB = hi(n); }
                const f2 w2 = mul2(am2, dla2);                             // w = opacity * Gs * dL/dalpha
                const f2 cr2 = mul2(dchan2, S.g0), cgn2 = mul2(dchan2, S.g1), cb2 = mul2(dchan2, S.g2);
                // Per-pair sums are raw moments of w = Gs * dL/dGs in the WHITENED offset (u, v) = L^T d, which the
                // exponent above has already formed (q = -(u^2 + v^2)):
                //   v = (S w u, S w v, S w u^2, S w u v, S w v^2, S w, dL/dr, dL/dg), d_b = dL/db
                // The per-Gaussian kernel turns them into dL/dpix = -(1/K) L (S w [u v]) and
                // dL/dcov2D = (1/2K^2) L (S w [u v][u v]^T) L^T, a congruence with the factor L it already holds.
                // Moments in pixel axes (S w dx^2, ...) say the same in exact arithmetic, but the way from them to
                // the covariance gradient divides by det^2 and subtracts terms that cancel for an elongated
                // Gaussian: at 256:1 axes the fp32 rounding of the accumulated moments came back 1e3 times larger
                // (scale / rotation gradients off by 4e-4, found by the referee checker).  |u|, |v| <= ~3.4 where
                // alpha >= 1/255, whatever the shape, so these sums are well scaled by construction.
                // (S w = opacity * S Gs dL/dalpha: the per-Gaussian kernel divides by the opacity for dL/dopacity)
                const f2 wu2 = mul2(w2, u2), wv2 = mul2(w2, v2);
                const f2 wuu2 = mul2(wu2, u2), wuv2 = mul2(wu2, v2), wvv2 = mul2(wv2, v2);
                float v[8];
                v[0] = lo(wu2) + hi(wu2);
                v[1] = lo(wv2) + hi(wv2);
                v[2] = lo(wuu2) + hi(wuu2);
                v[3] = lo(wuv2) + hi(wuv2);
                v[4] = lo(wvv2) + hi(wvv2);
                v[5] = lo(w2) + hi(w2);
                v[6] = lo(cr2) + hi(cr2);
                v[7] = lo(cgn2) + hi(cgn2);
                float d_b = lo(cb2) + hi(cb2);
                const float sum8 = warp_reduce8(v, lane);
                d_b = warp_sum(d_b);
                red_add_if(red_lane, acc + (size_t)__float_as_uint(f2v.w) * 12 + red_off, lane == 1 ? d_b : sum8);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// backward, variant "H4": one HALF warp per 8x8 block, four pixels per lane
// ------------------------------------------------------------------------------------------------
// The kernel above spends 42 of its 110 loop instructions reducing 9 sums over the 32 lanes of a warp, once per
// (8x8 block, Gaussian) hit.  Here a warp owns TWO 8x8 blocks (lanes 0-15 the left one, 16-31 the right one), each
// half walks its OWN hit list in lockstep with the other, and a lane carries four pixels of one column (two packed
// pairs): one loop iteration evaluates two hits, the reduction runs over 16 lanes (4 levels) and is paid once per
// iteration, i.e. half of it per hit, and the per-hit control (list pop, record fetch, atomic) is shared by the pair.
// The price: the halves wait for each other when their lists differ in length, and twice the pixel state per lane.
constexpr int H4_THREADS = 64;      // 2 warps x (2 blocks of 8x8) = one 16x16 tile
constexpr int H4_STAGE = BATCH / H4_THREADS;

struct Quad {          // the four pixels of a lane: column px, rows py0 + {0, 2, 4, 6}; pairs (0,2) and (4,6)
    f2 T01, T23, Sg01, Sg23;
    f2 g0_01, g0_23, g1_01, g1_23, g2_01, g2_23;
    unsigned int last0, last1, last2, last3;
};

// sum over the 16 lanes of a half warp: 8 values with a transposing butterfly (lane bits 3,2,1 pick the value the lane
// ends up holding, bit 0 duplicates), all shuffles stay inside the half (xor 8, 4, 2, 1)
__device__ __forceinline__ float half_reduce8(float v[8], int lane)
{
    bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float send = hi ? v[i] : v[i + 4];
        const float keep = hi ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = hi ? v[i] : v[i + 2];
        const float keep = hi ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    hi = lane & 2;
    {
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    return v[0];      // value index ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)
}
__device__ __forceinline__ float half_sum(float v)
{
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

__global__ void __launch_bounds__(H4_THREADS, 12)
render_backward_h4_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, unsigned long long cap,
                          const float* __restrict__ dL_dout, float* __restrict__ acc /* [P][12] */)
{
    __shared__ float4 s_feat[3 * BATCH];
    __shared__ unsigned char s_list[2][2][BATCH];
    __shared__ unsigned int s_max[2];

    pdl_prologue();
    const int Tv = s.gx * s.gy;
    const int view = (int)blockIdx.x / Tv;
    const int tile = (int)blockIdx.x - view * Tv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, l16 = lane & 15;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int tx = tile % s.gx, ty = tile / s.gx;
    const int by0 = ty * TILE + warp * 8;
    const int bx0 = tx * TILE + half * 8;                 // this half's block
    const int px = bx0 + (l16 & 7), py0 = by0 + (l16 >> 3);
    // both blocks of the warp, for the culling tests every lane makes for both halves
    BlockGeom gA, gB;
    gA.px = gB.px = 0; gA.py = gB.py = 0;
    gA.xmin = (float)(tx * TILE); gA.xmax = gA.xmin + 7.f; gB.xmin = gA.xmin + 8.f; gB.xmax = gB.xmin + 7.f;
    gA.ymin = gB.ymin = (float)by0; gA.ymax = gB.ymax = (float)(by0 + 7);
    const size_t N = (size_t)s.W * s.H;
    const float wgt = s.vt.weight[view];
    const float* const dL = dL_dout + (size_t)s.vt.out_image[view] * 3 * N;
    const float* const fT = im.final_T + (size_t)view * N;
    const unsigned int* const nc = im.n_contrib + (size_t)view * N;
    const int ox = s.vt.flip_x[view] ? s.W - 1 - px : px;

    const uint2 rg = im.ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0) return;
    if ((unsigned long long)rg.y > cap || im.hdr->overflow != 0u) return;

    unsigned int spec_id[H4_STAGE];
#pragma unroll
    for (int r = 0; r < H4_STAGE; r++) {
        const int kp = n - 1 - (tid + r * H4_THREADS);
        spec_id[r] = kp >= 0 ? __ldg(bin.point_list + rg.x + kp) : 0u;
    }

    const float bg0 = __ldg(s.bg), bg1 = __ldg(s.bg + 1), bg2 = __ldg(s.bg + 2);
    Quad S;
    {
        float T[4], Sg[4], g0[4], g1[4], g2[4];
        unsigned int last[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int py = py0 + 2 * i;
            const bool in = px < s.W && py < s.H;
            T[i] = 0.f; Sg[i] = 0.f; g0[i] = g1[i] = g2[i] = 0.f; last[i] = 0u;
            if (in) {
                const size_t pix = (size_t)py * s.W + px, opix = (size_t)py * s.W + ox;
                T[i] = fT[pix]; last[i] = nc[pix];
                g0[i] = wgt * dL[opix]; g1[i] = wgt * dL[N + opix]; g2[i] = wgt * dL[2 * N + opix];
                Sg[i] = T[i] * (bg0 * g0[i] + bg1 * g1[i] + bg2 * g2[i]);
            }
        }
        S.T01 = mk2(T[0], T[1]); S.T23 = mk2(T[2], T[3]); S.Sg01 = mk2(Sg[0], Sg[1]); S.Sg23 = mk2(Sg[2], Sg[3]);
        S.g0_01 = mk2(g0[0], g0[1]); S.g0_23 = mk2(g0[2], g0[3]); S.g1_01 = mk2(g1[0], g1[1]); S.g1_23 = mk2(g1[2], g1[3]);
        S.g2_01 = mk2(g2[0], g2[1]); S.g2_23 = mk2(g2[2], g2[3]);
        S.last0 = last[0]; S.last1 = last[1]; S.last2 = last[2]; S.last3 = last[3];
    }
    const float pxf = (float)px, pyf = (float)py0;
    const f2 npy01 = mk2(-pyf, -(pyf + 2.f)), npy23 = mk2(-(pyf + 4.f), -(pyf + 6.f));

    // deepest contributor of each block (the replay of a block starts there) and of the tile
    const unsigned int lmax = max(max(S.last0, S.last1), max(S.last2, S.last3));
    const unsigned int hmaxA = __reduce_max_sync(FULL, half == 0 ? lmax : 0u);
    const unsigned int hmaxB = __reduce_max_sync(FULL, half == 1 ? lmax : 0u);
    if (lane == 0) s_max[warp] = max(hmaxA, hmaxB);
    __syncthreads();
    const int m_len = min((int)max(s_max[0], s_max[1]), n);

    // lanes with bit 0 clear hold the 8 reduced sums of their half (index from lane bits 3, 2, 1), lane 1 of each half
    // the ninth
    const int red_lane = ((l16 & 1) == 0 || l16 == 1) ? 1 : 0;
    const int red_off = l16 == 1 ? 8 : (((l16 >> 3) & 1) * 4 + ((l16 >> 2) & 1) * 2 + ((l16 >> 1) & 1));

    for (int base = 0; base < m_len; base += BATCH) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < H4_STAGE; r++) {
            const int slot = tid + r * H4_THREADS;
            const int kpos = m_len - 1 - (base + slot);
            if (kpos >= 0) {
                const unsigned int id = (base == 0 && m_len == n) ? spec_id[r] : bin.point_list[rg.x + kpos];
                stage(s_feat, slot, geo, id);
            }
        }
        __syncthreads();
        const int cnt = min(BATCH, m_len - base);
        // the two halves' hit lists of the whole batch (slot numbers, back to front), so that the halves only wait for
        // each other once per batch, not once per chunk of 32 candidates
        int nA = 0, nB = 0;
        for (int c = 0; c < cnt; c += 32) {
            const int pos = m_len - 1 - (base + c) - lane;
            bool hitA = false, hitB = false;
            if (c + lane < cnt) {
                const float4* e = s_feat + 3 * (c + lane);
                if ((unsigned int)pos < hmaxA) hitA = block_hit(gA, e);
                if ((unsigned int)pos < hmaxB) hitB = block_hit(gB, e);
            }
            const unsigned int mA = __ballot_sync(FULL, hitA), mB = __ballot_sync(FULL, hitB);
            if (hitA) s_list[warp][0][nA + __popc(mA & lt_mask)] = (unsigned char)(c + lane);
            if (hitB) s_list[warp][1][nB + __popc(mB & lt_mask)] = (unsigned char)(c + lane);
            nA += __popc(mA); nB += __popc(mB);
        }
        __syncwarp();
        {
            const int nh = half ? nB : nA;
            const int iters = max(nA, nB);
            const unsigned char* const list = s_list[warp][half];
            const int posb = m_len - 1 - base;
            for (int it = 0; it < iters; it++) {
                const bool valid = it < nh;
                const int k = valid ? (int)list[it] : 0;
                const int pos0 = posb;
                const float4* e = s_feat + 3 * k;
                const float4 f0 = e[0];
                const float4 f1 = e[1];
                const float4 f2v = e[2];
                const float dx = f0.x - pxf;
                const f2 dy01 = add2(bc2(f0.y), npy01), dy23 = add2(bc2(f0.y), npy23);
                f2 u01, v01, u23, v23;
                const f2 nq01 = neg_falloff_log2(f0.z, f1, dx, dy01, u01, v01);
                const f2 nq23 = neg_falloff_log2(f0.z, f1, dx, dy23, u23, v23);
                const f2 a01 = mul2(bc2(f1.w), mk2(ex2_approx(-lo(nq01)), ex2_approx(-hi(nq01))));
                const f2 a23 = mul2(bc2(f1.w), mk2(ex2_approx(-lo(nq23)), ex2_approx(-hi(nq23))));
                const unsigned int pos = (unsigned int)(pos0 - k);
                const bool u0 = valid && (pos < S.last0) && !(lo(a01) < ALPHA_MIN);
                const bool u1 = valid && (pos < S.last1) && !(hi(a01) < ALPHA_MIN);
                const bool u2 = valid && (pos < S.last2) && !(lo(a23) < ALPHA_MIN);
                const bool u3 = valid && (pos < S.last3) && !(hi(a23) < ALPHA_MIN);
                const f2 am01 = mk2(u0 ? lo(a01) : 0.f, u1 ? hi(a01) : 0.f);
                const f2 am23 = mk2(u2 ? lo(a23) : 0.f, u3 ? hi(a23) : 0.f);
                const f2 ae01 = mk2(fminf(ALPHA_MAX, lo(am01)), fminf(ALPHA_MAX, hi(am01)));
                const f2 ae23 = mk2(fminf(ALPHA_MAX, lo(am23)), fminf(ALPHA_MAX, hi(am23)));
                const f2 om01 = sub2(bc2(1.f), ae01), om23 = sub2(bc2(1.f), ae23);
                const f2 ra01 = mk2(rcp_approx(lo(om01)), rcp_approx(hi(om01)));
                const f2 ra23 = mk2(rcp_approx(lo(om23)), rcp_approx(hi(om23)));
                S.T01 = mul2(S.T01, ra01); S.T23 = mul2(S.T23, ra23);
                const f2 cg01 = fma2(bc2(f2v.x), S.g0_01, fma2(bc2(f2v.y), S.g1_01, mul2(bc2(f2v.z), S.g2_01)));
                const f2 cg23 = fma2(bc2(f2v.x), S.g0_23, fma2(bc2(f2v.y), S.g1_23, mul2(bc2(f2v.z), S.g2_23)));
                const f2 dla01 = fma2(S.T01, cg01, neg2(mul2(ra01, S.Sg01)));
                const f2 dla23 = fma2(S.T23, cg23, neg2(mul2(ra23, S.Sg23)));
                const f2 dch01 = mul2(ae01, S.T01), dch23 = mul2(ae23, S.T23);
                S.Sg01 = fma2(dch01, cg01, S.Sg01); S.Sg23 = fma2(dch23, cg23, S.Sg23);
                const f2 w01 = mul2(am01, dla01), w23 = mul2(am23, dla23);
                const f2 wu01 = mul2(w01, u01), wu23 = mul2(w23, u23), wv01 = mul2(w01, v01), wv23 = mul2(w23, v23);
                // per-lane partial sums of the four pixels: pairs first (packed), then the two halves of the pair
                const f2 s_wu = add2(wu01, wu23), s_wv = add2(wv01, wv23);
                const f2 s_wuu = fma2(wu01, u01, mul2(wu23, u23)), s_wuv = fma2(wu01, v01, mul2(wu23, v23));
                const f2 s_wvv = fma2(wv01, v01, mul2(wv23, v23)), s_w = add2(w01, w23);
                const f2 s_r = fma2(dch01, S.g0_01, mul2(dch23, S.g0_23)), s_g = fma2(dch01, S.g1_01, mul2(dch23, S.g1_23));
                const f2 s_b = fma2(dch01, S.g2_01, mul2(dch23, S.g2_23));
                float v[8];
                v[0] = lo(s_wu) + hi(s_wu);
                v[1] = lo(s_wv) + hi(s_wv);
                v[2] = lo(s_wuu) + hi(s_wuu);
                v[3] = lo(s_wuv) + hi(s_wuv);
                v[4] = lo(s_wvv) + hi(s_wvv);
                v[5] = lo(s_w) + hi(s_w);
                v[6] = lo(s_r) + hi(s_r);
                v[7] = lo(s_g) + hi(s_g);
                float d_b = lo(s_b) + hi(s_b);
                const float sum8 = half_reduce8(v, lane);
                d_b = half_sum(d_b);
                // (an idle half — its list ran out — reduces zeros; its atomic is predicated off)
                red_add_if(valid ? red_lane : 0, acc + (size_t)__float_as_uint(f2v.w) * 12 + red_off, l16 == 1 ? d_b : sum8);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, variant "2P": two phases per warp — replay with lanes = pixels, moments with lanes = Gaussians
// ------------------------------------------------------------------------------------------------
// In the kernel above more than half of the loop (58 of 111 instructions per hit) sums 9 values over the 32 lanes.
// Only two numbers per pixel come out of the sequential part of the replay: w = opacity Gs dL/dalpha and
// dch = alpha T.  Phase 1 (lanes = pixel pairs, as above) stops there and parks (wA, wB, dchA, dchB) of the hit in a
// per-warp shared-memory row: one STS.128 per lane.  After HB hits, phase 2 turns the warp around: lane = (hit, part
// of the block's pixels), each lane walks its 64/PARTS pixels, rebuilds the whitened offset (u, v) of ITS Gaussian
// with the same expressions as phase 1 and accumulates the 9 sums in registers — no cross-lane traffic inside the
// loop, one butterfly over the PARTS lanes of a hit at the end and one set of atomics per hit.  Rows are padded to 33
// float4 so that both the row-wise writes of phase 1 and the column-wise reads of phase 2 are free of bank conflicts.
template <int HB>
struct TwoPhase {
    static constexpr int PARTS = 32 / HB;   // lanes per hit in phase 2
    static constexpr int ITER = HB;         // pixel pairs per lane in phase 2 (32 pairs / PARTS)
    static constexpr int ROW = 33;          // float4 per parked hit (32 lanes + 1 pad)
};

__device__ __forceinline__ void red_add_v4(float* a, float x, float y, float z, float w)
{
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int HB>
__device__ __forceinline__ void flush_hits(const float4* __restrict__ s_feat, const float4* __restrict__ wd,
                                           const float4* __restrict__ g01, const float2* __restrict__ g2,
                                           const int* __restrict__ hit, int cnt, int lane, float bx0f, float by0f,
                                           float* __restrict__ acc)
{
    using P2 = TwoPhase<HB>;
    const int h = lane & (HB - 1), part = lane / HB;
    const int slot = hit[h];                                   // (rows >= cnt hold an older, still valid slot)
    const float4 f0 = s_feat[3 * slot];
    const float4 f1 = s_feat[3 * slot + 1];
    const unsigned int id = __float_as_uint(s_feat[3 * slot + 2].w);
    const float4* const row = wd + h * P2::ROW + P2::ITER * part;
    const float4* const ga = g01 + (P2::ITER + 1) * part;      // (one pad entry per part: the parts' loads of the same
    const float2* const gb = g2 + (P2::ITER + 1) * part;       //  i fall on different banks)
    const float pyb = by0f + (float)((P2::ITER / 8) * part);   // first pixel row of this lane's part
    f2 a_wu = bc2(0.f), a_wv = bc2(0.f), a_wuu = bc2(0.f), a_wuv = bc2(0.f), a_wvv = bc2(0.f), a_w = bc2(0.f);
    f2 a_r = bc2(0.f), a_g = bc2(0.f), a_b = bc2(0.f);
#pragma unroll
    for (int i = 0; i < P2::ITER; i++) {
        const float4 q = row[i];                               // (wA, wB, dchA, dchB) of pixel pair ITER * part + i
        const float4 gq = ga[i];                               // (g0A, g0B, g1A, g1B)
        const float2 gr = gb[i];                               // (g2A, g2B)
        const float pxf = bx0f + (float)(i & 7);
        const float pyf = pyb + (float)(i >> 3);
        const float dx = f0.x - pxf;
        const f2 dy2 = add2(bc2(f0.y), mk2(-pyf, -(pyf + 4.f)));
        f2 u2, v2;
        (void)neg_falloff_log2(f0.z, f1, dx, dy2, u2, v2);
        const f2 w2 = mk2(q.x, q.y), dch2 = mk2(q.z, q.w);
        const f2 wu2 = mul2(w2, u2), wv2 = mul2(w2, v2);
        a_wu = add2(a_wu, wu2);
        a_wv = add2(a_wv, wv2);
        fma2_acc(a_wuu, wu2, u2);
        fma2_acc(a_wuv, wu2, v2);
        fma2_acc(a_wvv, wv2, v2);
        a_w = add2(a_w, w2);
        fma2_acc(a_r, dch2, mk2(gq.x, gq.y));
        fma2_acc(a_g, dch2, mk2(gq.z, gq.w));
        fma2_acc(a_b, dch2, mk2(gr.x, gr.y));
    }
    float v[9];
    v[0] = lo(a_wu) + hi(a_wu);
    v[1] = lo(a_wv) + hi(a_wv);
    v[2] = lo(a_wuu) + hi(a_wuu);
    v[3] = lo(a_wuv) + hi(a_wuv);
    v[4] = lo(a_wvv) + hi(a_wvv);
    v[5] = lo(a_w) + hi(a_w);
    v[6] = lo(a_r) + hi(a_r);
    v[7] = lo(a_g) + hi(a_g);
    v[8] = lo(a_b) + hi(a_b);
#pragma unroll
    for (int d = HB; d < 32; d <<= 1)
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] += __shfl_xor_sync(FULL, v[k], d);
    if (part == 0 && h < cnt) {
        float* const a = acc + (size_t)id * 12;                // 48-byte rows: two 16-byte vector reductions + a scalar
        red_add_v4(a, v[0], v[1], v[2], v[3]);
        red_add_v4(a + 4, v[4], v[5], v[6], v[7]);
        atomicAdd(a + 8, v[8]);
    }
}

template <int HB, int NB, int MINB>
__global__ void __launch_bounds__(BLEND_THREADS, MINB)
render_backward_2p_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, unsigned long long cap,
                          const float* __restrict__ dL_dout, float* __restrict__ acc /* [P][12] */)
{
    using P2 = TwoPhase<HB>;
    __shared__ float4 s_feat[3 * NB];
    __shared__ float4 s_wd[BLEND_WARPS][HB * P2::ROW];
    __shared__ float4 s_g01[BLEND_WARPS][32 + P2::PARTS];
    __shared__ float2 s_g2[BLEND_WARPS][32 + P2::PARTS];
    __shared__ int s_hit[BLEND_WARPS][HB];
    __shared__ unsigned int s_max[BLEND_WARPS];

    pdl_prologue();
    const int Tv = s.gx * s.gy;
    const int view = (int)blockIdx.x / Tv;          // virtual tile = view * Tv + tile
    const int tile = (int)blockIdx.x - view * Tv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BlockGeom bg = block_geom(tile, s.gx, tid);
    const bool inA = bg.px < s.W && bg.py < s.H, inB = bg.px < s.W && bg.py + 4 < s.H;
    const float pxf = (float)bg.px, pyf = (float)bg.py;
    const size_t N = (size_t)s.W * s.H;
    const size_t pixA = (size_t)bg.py * s.W + bg.px, pixB = (size_t)(bg.py + 4) * s.W + bg.px;
    const int ox = s.vt.flip_x[view] ? s.W - 1 - bg.px : bg.px;
    const size_t opixA = (size_t)bg.py * s.W + ox, opixB = (size_t)(bg.py + 4) * s.W + ox;
    const float wgt = s.vt.weight[view];
    const float* const dL = dL_dout + (size_t)s.vt.out_image[view] * 3 * N;
    const float* const fT = im.final_T + (size_t)view * N;
    const unsigned int* const nc = im.n_contrib + (size_t)view * N;

    const uint2 rg = im.ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0) return;
    if ((unsigned long long)rg.y > cap || im.hdr->overflow != 0u) return;   // see render_backward_kernel

    unsigned int spec_id[NB / BLEND_THREADS];
#pragma unroll
    for (int r = 0; r < NB / BLEND_THREADS; r++) {
        const int kp = n - 1 - (tid + r * BLEND_THREADS);
        spec_id[r] = kp >= 0 ? __ldg(bin.point_list + rg.x + kp) : 0u;
    }

    const float bg0 = __ldg(s.bg), bg1 = __ldg(s.bg + 1), bg2 = __ldg(s.bg + 2);
    PairBwd S{};
    {
        float gA0 = 0.f, gA1 = 0.f, gA2 = 0.f, gB0 = 0.f, gB1 = 0.f, gB2 = 0.f;
        if (inA) {
            S.TA = fT[pixA]; S.lastA = nc[pixA];
            gA0 = wgt * dL[opixA]; gA1 = wgt * dL[N + opixA]; gA2 = wgt * dL[2 * N + opixA];
            S.SgA = S.TA * (bg0 * gA0 + bg1 * gA1 + bg2 * gA2);
        }
        if (inB) {
            S.TB = fT[pixB]; S.lastB = nc[pixB];
            gB0 = wgt * dL[opixB]; gB1 = wgt * dL[N + opixB]; gB2 = wgt * dL[2 * N + opixB];
            S.SgB = S.TB * (bg0 * gB0 + bg1 * gB1 + bg2 * gB2);
        }
        S.g0 = mk2(gA0, gB0); S.g1 = mk2(gA1, gB1); S.g2 = mk2(gA2, gB2);
        s_g01[warp][lane + lane / P2::ITER] = make_float4(gA0, gB0, gA1, gB1);   // phase 2 reads the pixels' dL/dC here
        s_g2[warp][lane + lane / P2::ITER] = make_float2(gA2, gB2);
        if (lane < HB) s_hit[warp][lane] = 0;
    }
    const f2 npy2 = mk2(-pyf, -(pyf + 4.f));
    const float bx0f = (float)(bg.px - (lane & 7)), by0f = (float)(bg.py - (lane >> 3));

    const unsigned int wmax = __reduce_max_sync(FULL, max(S.lastA, S.lastB));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    unsigned int bmax = 0;
#pragma unroll
    for (int w = 0; w < BLEND_WARPS; w++) bmax = max(bmax, s_max[w]);
    const int m_len = min((int)bmax, n);

    float4* const my_wd = s_wd[warp];
    int parked = 0;                                   // hits parked in my_wd (warp-uniform)

    for (int base = 0; base < m_len; base += NB) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < NB / BLEND_THREADS; r++) {
            const int slot = tid + r * BLEND_THREADS;
            const int kpos = m_len - 1 - (base + slot);
            if (kpos >= 0) {
                const unsigned int id = (base == 0 && m_len == n) ? spec_id[r] : bin.point_list[rg.x + kpos];
                stage(s_feat, slot, geo, id);
            }
        }
        __syncthreads();
        const int cnt = min(NB, m_len - base);
        for (int c = 0; c < cnt; c += 32) {
            const int pos0 = m_len - 1 - (base + c);
            bool hit = false;
            if (c + lane < cnt && (unsigned int)(pos0 - lane) < wmax) hit = block_hit(bg, s_feat + 3 * (c + lane));
            unsigned int m = __ballot_sync(FULL, hit);
            const float4* chunk = s_feat + 3 * c;
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const float4* e = chunk + 3 * k;
                const float4 f0 = e[0];
                const float4 f1 = e[1];
                const float4 f2v = e[2];
                const float dx = f0.x - pxf;
                const f2 dy2 = add2(bc2(f0.y), npy2);
                const f2 nq2 = neg_falloff_log2(f0.z, f1, dx, dy2);
                const f2 Gs2 = mk2(ex2_approx(-lo(nq2)), ex2_approx(-hi(nq2)));
                const f2 a2 = mul2(bc2(f1.w), Gs2);
                const unsigned int pos = (unsigned int)(pos0 - k);
                const bool useA = (pos < S.lastA) && !(lo(a2) < ALPHA_MIN);
                const bool useB = (pos < S.lastB) && !(hi(a2) < ALPHA_MIN);
                const f2 am2 = mk2(useA ? lo(a2) : 0.f, useB ? hi(a2) : 0.f);
                const f2 ae2 = mk2(fminf(ALPHA_MAX, lo(am2)), fminf(ALPHA_MAX, hi(am2)));
                const f2 om2 = sub2(bc2(1.f), ae2);
                const f2 ra2 = mk2(rcp_approx(lo(om2)), rcp_approx(hi(om2)));
                const f2 T2 = mul2(mk2(S.TA, S.TB), ra2);
                S.TA = lo(T2); S.TB = hi(T2);
                const f2 cg2 = fma2(bc2(f2v.x), S.g0, fma2(bc2(f2v.y), S.g1, mul2(bc2(f2v.z), S.g2)));
                const f2 Sg2 = mk2(S.SgA, S.SgB);
                const f2 dla2 = fma2(T2, cg2, neg2(mul2(ra2, Sg2)));
                const f2 dchan2 = mul2(ae2, T2);
                { const f2 nn = fma2(dchan2, cg2, Sg2); S.SgA = lo(nn); S.SgB = hi(nn); }
                const f2 w2 = mul2(am2, dla2);
                my_wd[parked * P2::ROW + lane] = make_float4(lo(w2), hi(w2), lo(dchan2), hi(dchan2));
                if (lane == 0) s_hit[warp][parked] = c + k;
                parked++;
                if (parked == HB) {
                    __syncwarp();
                    flush_hits<HB>(s_feat, my_wd, s_g01[warp], s_g2[warp], s_hit[warp], HB, lane, bx0f, by0f, acc);
                    __syncwarp();
                    parked = 0;
                }
            }
        }
        if (parked) {     // the staged records are about to be replaced
            __syncwarp();
            flush_hits<HB>(s_feat, my_wd, s_g01[warp], s_g2[warp], s_hit[warp], parked, lane, bx0f, by0f, acc);
            __syncwarp();
            parked = 0;
        }
    }
}

cudaError_t launch_render_backward(const DevSettings& s, int P, GeomView g, ImageView im, BinView b, long long cap,
                                   const float* dL_dout, float4* acc, bool acc_is_zero, cudaStream_t st)
{
    if (!acc_is_zero) {
        cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)P * s.n_views * 48, st);
        if (e != cudaSuccess) return e;
    }
    const int T = s.gx * s.gy * s.n_views;
    if (T <= 0 || P <= 0) return cudaSuccess;
    count_launch();
    static const int variant = [] { const char* e = getenv("GSVC_BWD_VARIANT"); return e ? atoi(e) : 0; }();
    if (variant == 1)
        return launch_pdl(render_backward_h4_kernel, dim3(T), dim3(H4_THREADS), st, s, g, im, b, (unsigned long long)cap,
                          dL_dout, reinterpret_cast<float*>(acc));
    if (variant == 2)
        return launch_pdl(render_backward_2p_kernel<8, 256, 6>, dim3(T), dim3(BLEND_THREADS), st, s, g, im, b,
                          (unsigned long long)cap, dL_dout, reinterpret_cast<float*>(acc));
    if (variant == 3)
        return launch_pdl(render_backward_2p_kernel<16, 128, 5>, dim3(T), dim3(BLEND_THREADS), st, s, g, im, b,
                          (unsigned long long)cap, dL_dout, reinterpret_cast<float*>(acc));
    return launch_pdl(render_backward_kernel, dim3(T), dim3(BLEND_THREADS), st, s, g, im, b, (unsigned long long)cap,
                      dL_dout, reinterpret_cast<float*>(acc));
}

}  // namespace gsvc
