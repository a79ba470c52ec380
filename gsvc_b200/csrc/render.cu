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
#include "common.cuh"

namespace gsvc {

constexpr int BLEND_THREADS = 128;  // 4 warps x (8x8 pixels), 2 pixels per lane
constexpr int BLEND_WARPS = BLEND_THREADS / 32;
constexpr int BATCH = 256;          // Gaussians staged per round (2 per thread)
constexpr unsigned FULL = 0xffffffffu;
constexpr float LOG2E = 1.4426950408889634f;

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
//   [0] = (pix.x, pix.y, B/A, B/C)
//   [1] = (A', B', C', opacity) with A' = -A log2e/2, B' = -B log2e, C' = -C log2e/2, so that
//         alpha = opacity * 2^q(d),  q(d) = A' dx^2 + B' dx dy + C' dy^2  (q <= 0, concave)
//   [2] = (r, g, b, thr) with thr = -log2(255 opacity) - margin: alpha >= 1/255  <=>  q >= thr (+margin)
constexpr float CULL_MARGIN = 1e-3f;  // in log2 units (7e-4 relative in alpha) >> fp32 rounding of q

__device__ __forceinline__ void stage(float4* s_feat, int slot, const GeomView& geo, unsigned int id)
{
    float4 f0 = __ldg(geo.feat0 + id);
    float4 f1 = __ldg(geo.feat1 + id);
    float4 f2 = __ldg(geo.feat2 + id);
    f0.z = __fdividef(f1.y, f1.x);
    f0.w = __fdividef(f1.y, f1.z);
    f2.w = -__log2f(255.0f * f1.w) - CULL_MARGIN;
    f1.x *= -0.5f * LOG2E;
    f1.y *= -LOG2E;
    f1.z *= -0.5f * LOG2E;
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
    const float thr = e[2].w;
    const float ex = fminf(fmaxf(f0.x, b.xmin), b.xmax), ey = fminf(fmaxf(f0.y, b.ymin), b.ymax);
    const float dxe = ex - f0.x, dye = ey - f0.y;
    const float dy1 = fminf(fmaxf(fmaf(-f0.w, dxe, f0.y), b.ymin), b.ymax) - f0.y;
    const float dx2 = fminf(fmaxf(fmaf(-f0.z, dye, f0.x), b.xmin), b.xmax) - f0.x;
    const float q1 = fmaf(fmaf(f1.x, dxe, f1.y * dy1), dxe, (f1.z * dy1) * dy1);
    const float q2 = fmaf(fmaf(f1.x, dx2, f1.y * dye), dx2, (f1.z * dye) * dye);
    return fmaxf(q1, q2) >= thr;
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
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const BlockGeom bg = block_geom(tile, s.gx, tid);
    const bool inA = bg.px < s.W && bg.py < s.H, inB = bg.px < s.W && bg.py + 4 < s.H;
    const float pxf = (float)bg.px, pyf = (float)bg.py;

    const uint2 rg = im.ranges[tile];
    if ((unsigned long long)rg.y > cap) return;  // capacity overflow: host re-runs with a larger buffer
    const int n = (int)(rg.y - rg.x);

    float TA = 1.f, A0 = 0.f, A1 = 0.f, A2 = 0.f;   // pixel A = (px, py)
    float TB = 1.f, B0 = 0.f, B1 = 0.f, B2 = 0.f;   // pixel B = (px, py + 4)
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
                const float4 f2 = e[2];
                const float dx = f0.x - pxf, dyA = f0.y - pyf, dyB = dyA - 4.f;
                const float t = f1.x * dx;                                          // A' dx
                const float qA = fmaf(fmaf(f1.y, dyA, t), dx, (f1.z * dyA) * dyA);  // log2 of the falloff
                const float qB = fmaf(fmaf(f1.y, dyB, t), dx, (f1.z * dyB) * dyB);
                const float aA = fminf(ALPHA_MAX, f1.w * ex2_approx(qA));
                const float aB = fminf(ALPHA_MAX, f1.w * ex2_approx(qB));
                // the reference `continue`s on power > 0 and on alpha < 1/255 (and never gets here once done)
                const bool okA = !(qA > 0.f) && !(aA < minA);
                const bool okB = !(qB > 0.f) && !(aB < minB);
                const unsigned int pos = pos1 + (unsigned int)k;
                const float tTA = TA * (1.f - aA), tTB = TB * (1.f - aB);
                // T' < 1e-4: the pixel stops and this Gaussian is NOT added
                if (okA && tTA < T_STOP) minA = DONE;
                if (okB && tTB < T_STOP) minB = DONE;
                if (okA && !(tTA < T_STOP)) {
                    const float w = aA * TA;
                    A0 = fmaf(f2.x, w, A0); A1 = fmaf(f2.y, w, A1); A2 = fmaf(f2.z, w, A2);
                    TA = tTA;
                    lastA = pos;
                }
                if (okB && !(tTB < T_STOP)) {
                    const float w = aB * TB;
                    B0 = fmaf(f2.x, w, B0); B1 = fmaf(f2.y, w, B1); B2 = fmaf(f2.z, w, B2);
                    TB = tTB;
                    lastB = pos;
                }
            }
        }
    }
    const size_t N = (size_t)s.W * s.H;
    const float bg0 = __ldg(s.bg + 0), bg1 = __ldg(s.bg + 1), bg2 = __ldg(s.bg + 2);
    if (inA) {
        const size_t pix = (size_t)bg.py * s.W + bg.px;
        out_color[pix] = A0 + TA * bg0;
        out_color[N + pix] = A1 + TA * bg1;
        out_color[2 * N + pix] = A2 + TA * bg2;
        im.final_T[pix] = TA;
        im.n_contrib[pix] = lastA;
    }
    if (inB) {
        const size_t pix = (size_t)(bg.py + 4) * s.W + bg.px;
        out_color[pix] = B0 + TB * bg0;
        out_color[N + pix] = B1 + TB * bg1;
        out_color[2 * N + pix] = B2 + TB * bg2;
        im.final_T[pix] = TB;
        im.n_contrib[pix] = lastB;
    }
}

cudaError_t launch_render_forward(const DevSettings& s, GeomView g, ImageView im, BinView b, long long cap,
                                  float* out_color, cudaStream_t st)
{
    const int T = s.gx * s.gy;
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

// Per-pixel replay state of the backward.
struct PixBwd {
    float T;                   // running transmittance (T_i before the current Gaussian)
    float Sg;                  // g . (colour composited behind the current Gaussian, incl. T_final * bg):
                               //   Sg_i = sum_{j>i} (c_j . g) alpha_j T_j + T_final (bg . g)
    float g0, g1, g2;          // dL/dC of the pixel
    unsigned int last;         // n_contrib
};

// One (pixel, Gaussian) pair, branch-free: a pair that does not contribute is blended with alpha = 0,
// which leaves T, the behind-colour sum and every gradient sum unchanged (bit-identically).
//   dC/dalpha_i = c_i T_i - (sum_{j>i} c_j alpha_j T_j + T_final bg) / (1 - alpha_i)
// contracted with g = dL/dC up front, so the behind-colour recursion is ONE scalar per pixel (Sg) instead of
// three colour channels.  Returns w = opacity * Gs * dL/dalpha (the weight of the conic / position moments);
// adds the colour and opacity terms to (s_op, s_r, s_g, s_b).
__device__ __forceinline__ float pair_backward(PixBwd& p, bool use, float alpha, float Gs, float opacity,
                                               const float4 f2, float& s_op, float& s_r, float& s_g, float& s_b)
{
    const float ae = use ? alpha : 0.f;
    const float ra = rcp_approx(1.f - ae);
    p.T *= ra;
    const float cg = fmaf(f2.x, p.g0, fmaf(f2.y, p.g1, f2.z * p.g2));
    float dla = fmaf(p.T, cg, -(ra * p.Sg));
    dla = use ? dla : 0.f;                 // U4: straight-through the 0.99 cap otherwise
    const float dchan = ae * p.T;
    s_r = fmaf(dchan, p.g0, s_r);
    s_g = fmaf(dchan, p.g1, s_g);
    s_b = fmaf(dchan, p.g2, s_b);
    p.Sg = fmaf(dchan, cg, p.Sg);
    const float gd = Gs * dla;
    s_op += gd;
    return opacity * gd;
}

__global__ void __launch_bounds__(BLEND_THREADS)
render_backward_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, const float* __restrict__ dL_dout,
                       float* __restrict__ acc /* [P][12] */)
{
    __shared__ float4 s_feat[3 * BATCH];
    __shared__ unsigned int s_id[BATCH];
    __shared__ unsigned int s_max[BLEND_WARPS];

    pdl_prologue();
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const BlockGeom bg = block_geom(tile, s.gx, tid);
    const bool inA = bg.px < s.W && bg.py < s.H, inB = bg.px < s.W && bg.py + 4 < s.H;
    const float pxf = (float)bg.px, pyf = (float)bg.py;
    const size_t N = (size_t)s.W * s.H;
    const size_t pixA = (size_t)bg.py * s.W + bg.px, pixB = (size_t)(bg.py + 4) * s.W + bg.px;

    const uint2 rg = im.ranges[tile];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0) return;

    const float bg0 = __ldg(s.bg), bg1 = __ldg(s.bg + 1), bg2 = __ldg(s.bg + 2);
    PixBwd A{}, B{};
    if (inA) {
        A.T = im.final_T[pixA]; A.last = im.n_contrib[pixA];
        A.g0 = dL_dout[pixA]; A.g1 = dL_dout[N + pixA]; A.g2 = dL_dout[2 * N + pixA];
        A.Sg = A.T * (bg0 * A.g0 + bg1 * A.g1 + bg2 * A.g2);
    }
    if (inB) {
        B.T = im.final_T[pixB]; B.last = im.n_contrib[pixB];
        B.g0 = dL_dout[pixB]; B.g1 = dL_dout[N + pixB]; B.g2 = dL_dout[2 * N + pixB];
        B.Sg = B.T * (bg0 * B.g0 + bg1 * B.g1 + bg2 * B.g2);
    }

    // nothing behind the deepest contributor of the tile matters: start the replay there
    const unsigned int wmax = __reduce_max_sync(FULL, max(A.last, B.last));
    if (lane == 0) s_max[tid >> 5] = wmax;
    __syncthreads();
    unsigned int bmax = 0;
#pragma unroll
    for (int w = 0; w < BLEND_WARPS; w++) bmax = max(bmax, s_max[w]);
    const int m_len = (int)bmax;  // list entries [0, m_len) are replayed, back to front

    // lanes 0,4,..,28 own the 8 reduced sums (index lane/4), lane 1 the ninth: one atomic instruction
    const bool red_lane = (lane & 3) == 0 || lane == 1;
    const int red_off = lane == 1 ? 8 : (lane >> 2);

    for (int base = 0; base < m_len; base += BATCH) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < BATCH / BLEND_THREADS; r++) {
            const int slot = tid + r * BLEND_THREADS;
            const int kpos = m_len - 1 - (base + slot);  // list position staged in this slot
            if (kpos >= 0) {
                const unsigned int id = bin.point_list[rg.x + kpos];
                s_id[slot] = id;
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
                const float dx = f0.x - pxf, dyA = f0.y - pyf, dyB = dyA - 4.f;
                const float t = f1.x * dx;
                const float qA = fmaf(fmaf(f1.y, dyA, t), dx, (f1.z * dyA) * dyA);
                const float qB = fmaf(fmaf(f1.y, dyB, t), dx, (f1.z * dyB) * dyB);
                const float GsA = ex2_approx(qA), GsB = ex2_approx(qB);
                const float aA = fminf(ALPHA_MAX, f1.w * GsA), aB = fminf(ALPHA_MAX, f1.w * GsB);
                const unsigned int pos = (unsigned int)(pos0 - k);
                const bool useA = (pos < A.last) && !(qA > 0.f) && !(aA < ALPHA_MIN);
                const bool useB = (pos < B.last) && !(qB > 0.f) && !(aB < ALPHA_MIN);
                // (no warp vote here: after the exact culling above 99.9 % of the evaluations that reach this point
                //  contribute to at least one pixel — ncu source counters — and a pair that does not adds zeros)

                // Per-pair sums are raw moments of w = Gs * dL/dGs; the per-Gaussian kernel turns them into
                // dL/dpix and dL/dconic (it knows A,B,C), which keeps ~9 FP32 ops out of this loop:
                //   v = (S w dx, S w dy, S w dx^2, S w dx dy, S w dy^2, S Gs dL/dalpha, dL/dr, dL/dg), d_b = dL/db
                const float4 f2 = e[2];
                float v[8];
                float d_b = 0.f;
                v[5] = 0.f; v[6] = 0.f; v[7] = 0.f;
                const float wA = pair_backward(A, useA, aA, GsA, f1.w, f2, v[5], v[6], v[7], d_b);
                const float wB = pair_backward(B, useB, aB, GsB, f1.w, f2, v[5], v[6], v[7], d_b);
                const float wyA = wA * dyA, wyB = wB * dyB;
                v[0] = (wA + wB) * dx;
                v[1] = wyA + wyB;
                v[2] = v[0] * dx;
                v[3] = v[1] * dx;
                v[4] = fmaf(wyA, dyA, wyB * dyB);
                const float sum8 = warp_reduce8(v, lane);
                d_b = warp_sum(d_b);
                if (red_lane) atomicAdd(acc + (size_t)s_id[c + k] * 12 + red_off, lane == 1 ? d_b : sum8);
            }
        }
    }
}

cudaError_t launch_render_backward(const DevSettings& s, int P, GeomView g, ImageView im, BinView b,
                                   const float* dL_dout, float4* acc, bool acc_is_zero, cudaStream_t st)
{
    if (!acc_is_zero) {
        cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)P * 48, st);
        if (e != cudaSuccess) return e;
    }
    const int T = s.gx * s.gy;
    if (T <= 0 || P <= 0) return cudaSuccess;
    count_launch();
    return launch_pdl(render_backward_kernel, dim3(T), dim3(BLEND_THREADS), st, s, g, im, b, dL_dout,
                      reinterpret_cast<float*>(acc));
}

}  // namespace gsvc
