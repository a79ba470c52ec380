// Kernels (3) and (4): front-to-back alpha blend forward and its backward
// (the reference reaches them through
//  /root/reference/ortho_gaussian_renderer/renderer.py:90-98 and loss.backward(), pipeline/train.py:462).
//
// One CTA per 16x16 tile, one thread per pixel.  Gaussians of the tile's depth-sorted list are
// gathered with float4 loads into shared memory in batches of 256 and broadcast to all pixels.
#include "common.cuh"

namespace gsvc {

constexpr int BLEND_THREADS = TILE_PIX;  // 256

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLEND_THREADS)
render_forward_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, unsigned long long cap,
                      float* __restrict__ out_color)
{
    __shared__ float4 s_f0[BLEND_THREADS];
    __shared__ float4 s_f1[BLEND_THREADS];
    __shared__ float s_b[BLEND_THREADS];

    const int tile = blockIdx.x;
    const int tx = tile % s.gx, ty = tile / s.gx;
    const int tid = threadIdx.x;
    const int px = tx * TILE + (tid & (TILE - 1)), py = ty * TILE + (tid >> 4);
    const bool inside = px < s.W && py < s.H;
    const float pxf = (float)px, pyf = (float)py;

    const uint2 rg = im.ranges[tile];
    if ((unsigned long long)rg.y > cap) return;  // capacity overflow: host re-runs with a larger buffer
    const int n = (int)(rg.y - rg.x);

    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    unsigned int contributor = 0, last = 0;
    bool done = !inside;

    for (int base = 0; base < n; base += BLEND_THREADS) {
        // block-wide vote doubles as the barrier that protects the staging buffers
        if (__syncthreads_count(done) == BLEND_THREADS) break;
        const int idx = base + tid;
        if (idx < n) {
            const unsigned int id = bin.point_list[rg.x + idx];
            s_f0[tid] = __ldg(geo.feat0 + id);
            s_f1[tid] = __ldg(geo.feat1 + id);
            s_b[tid] = __ldg(&geo.feat2[id].x);
        }
        __syncthreads();
        const int cnt = min(BLEND_THREADS, n - base);
        // warp-ballot early termination: a warp whose 32 pixels are all saturated skips the batch
        if (__ballot_sync(0xffffffffu, !done) == 0u) continue;
        for (int j = 0; !done && j < cnt; j++) {
            contributor++;
            const float4 f0 = s_f0[j];
            const float4 f1 = s_f1[j];
            const float dx = f0.x - pxf, dy = f0.y - pyf;
            const float power = -0.5f * (f0.z * dx * dx + f1.x * dy * dy) - f0.w * dx * dy;
            if (power > 0.f) continue;
            const float alpha = fminf(ALPHA_MAX, f1.y * __expf(power));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < T_STOP) { done = true; continue; }
            const float w = alpha * T;
            C0 += f1.z * w;
            C1 += f1.w * w;
            C2 += s_b[j] * w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t N = (size_t)s.W * s.H, pix = (size_t)py * s.W + px;
        out_color[pix] = C0 + T * __ldg(s.bg + 0);
        out_color[N + pix] = C1 + T * __ldg(s.bg + 1);
        out_color[2 * N + pix] = C2 + T * __ldg(s.bg + 2);
        im.final_T[pix] = T;
        im.n_contrib[pix] = last;
    }
}

cudaError_t launch_render_forward(const DevSettings& s, GeomView g, ImageView im, BinView b, long long cap,
                                  float* out_color, cudaStream_t st)
{
    const int T = s.gx * s.gy;
    if (T <= 0) return cudaSuccess;
    render_forward_kernel<<<T, BLEND_THREADS, 0, st>>>(s, g, im, b, (unsigned long long)cap, out_color);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// backward: per-pixel replay back-to-front; per-Gaussian gradients are reduced across the warp with
// shuffles before a single lane touches global memory with atomics.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(BLEND_THREADS)
render_backward_kernel(DevSettings s, GeomView geo, ImageView im, BinView bin, const float* __restrict__ dL_dout,
                       float* __restrict__ acc /* [P][12] */)
{
    __shared__ float4 s_f0[BLEND_THREADS];
    __shared__ float4 s_f1[BLEND_THREADS];
    __shared__ float s_b[BLEND_THREADS];
    __shared__ unsigned int s_id[BLEND_THREADS];
    __shared__ unsigned int s_max[BLEND_THREADS / 32];

    const int tile = blockIdx.x;
    const int tx = tile % s.gx, ty = tile / s.gx;
    const int tid = threadIdx.x, lane = tid & 31;
    const int px = tx * TILE + (tid & (TILE - 1)), py = ty * TILE + (tid >> 4);
    const bool inside = px < s.W && py < s.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t N = (size_t)s.W * s.H, pix = (size_t)py * s.W + px;

    const uint2 rg = im.ranges[tile];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0) return;

    const float T_final = inside ? im.final_T[pix] : 0.f;
    const unsigned int last = inside ? im.n_contrib[pix] : 0u;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (inside) { g0 = dL_dout[pix]; g1 = dL_dout[N + pix]; g2 = dL_dout[2 * N + pix]; }
    const float bg_dot = __ldg(s.bg) * g0 + __ldg(s.bg + 1) * g1 + __ldg(s.bg + 2) * g2;

    // nothing behind the deepest contributor of the tile matters: start the replay there
    const unsigned int wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) s_max[tid >> 5] = wmax;
    __syncthreads();
    unsigned int bmax = 0;
#pragma unroll
    for (int w = 0; w < BLEND_THREADS / 32; w++) bmax = max(bmax, s_max[w]);
    const int m = (int)bmax;  // list entries [0, m) are replayed, back to front

    float T = T_final;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;      // colour composited behind the current Gaussian
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;   // previous (deeper) contributor's colour and alpha
    float last_alpha = 0.f;

    for (int base = 0; base < m; base += BLEND_THREADS) {
        __syncthreads();
        const int k = m - 1 - (base + tid);  // list position staged by this thread
        if (k >= 0) {
            const unsigned int id = bin.point_list[rg.x + k];
            s_id[tid] = id;
            s_f0[tid] = __ldg(geo.feat0 + id);
            s_f1[tid] = __ldg(geo.feat1 + id);
            s_b[tid] = __ldg(&geo.feat2[id].x);
        }
        __syncthreads();
        const int cnt = min(BLEND_THREADS, m - base);
        for (int j = 0; j < cnt; j++) {
            const unsigned int pos = (unsigned int)(m - 1 - (base + j));  // 0-based list position
            if (pos >= wmax) continue;                                     // warp-uniform
            const float4 f0 = s_f0[j];
            const float4 f1 = s_f1[j];
            const float cb = s_b[j];
            const float dx = f0.x - pxf, dy = f0.y - pyf;
            const float power = -0.5f * (f0.z * dx * dx + f1.x * dy * dy) - f0.w * dx * dy;
            const float Gs = __expf(power);
            const float alpha = fminf(ALPHA_MAX, f1.y * Gs);
            const bool use = (pos < last) && !(power > 0.f) && !(alpha < ALPHA_MIN);
            if (!__any_sync(0xffffffffu, use)) continue;

            float d_px = 0.f, d_py = 0.f, d_A = 0.f, d_B = 0.f, d_C = 0.f, d_op = 0.f, d_r = 0.f, d_g = 0.f, d_b = 0.f;
            if (use) {
                T = T / (1.f - alpha);
                const float dchan = alpha * T;
                a0 = last_alpha * lc0 + (1.f - last_alpha) * a0;
                a1 = last_alpha * lc1 + (1.f - last_alpha) * a1;
                a2 = last_alpha * lc2 + (1.f - last_alpha) * a2;
                lc0 = f1.z; lc1 = f1.w; lc2 = cb;
                float dL_dalpha = (f1.z - a0) * g0 + (f1.w - a1) * g1 + (cb - a2) * g2;
                d_r = dchan * g0; d_g = dchan * g1; d_b = dchan * g2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = f1.y * dL_dalpha;  // U4: straight-through the 0.99 cap
                const float gdx = Gs * dx, gdy = Gs * dy;
                d_px = dL_dG * (-gdx * f0.z - gdy * f0.w);
                d_py = dL_dG * (-gdy * f1.x - gdx * f0.w);
                d_A = -0.5f * gdx * dx * dL_dG;
                d_B = -gdx * dy * dL_dG;
                d_C = -0.5f * gdy * dy * dL_dG;
                d_op = Gs * dL_dalpha;
            }
            d_px = warp_sum(d_px); d_py = warp_sum(d_py);
            d_A = warp_sum(d_A); d_B = warp_sum(d_B); d_C = warp_sum(d_C);
            d_op = warp_sum(d_op);
            d_r = warp_sum(d_r); d_g = warp_sum(d_g); d_b = warp_sum(d_b);
            if (lane == 0) {
                float* a = acc + (size_t)s_id[j] * 12;
                atomicAdd(a + 0, d_px); atomicAdd(a + 1, d_py); atomicAdd(a + 2, d_A); atomicAdd(a + 3, d_B);
                atomicAdd(a + 4, d_C); atomicAdd(a + 5, d_op); atomicAdd(a + 6, d_r); atomicAdd(a + 7, d_g);
                atomicAdd(a + 8, d_b);
            }
        }
    }
}

cudaError_t launch_render_backward(const DevSettings& s, int P, GeomView g, ImageView im, BinView b,
                                   const float* dL_dout, float4* acc, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)P * 48, st);
    if (e != cudaSuccess) return e;
    const int T = s.gx * s.gy;
    if (T <= 0 || P <= 0) return cudaSuccess;
    render_backward_kernel<<<T, BLEND_THREADS, 0, st>>>(s, g, im, b, dL_dout, reinterpret_cast<float*>(acc));
    count_launch();
    return cudaGetLastError();
}

}  // namespace gsvc
