// Kernel group (2): tile/depth key duplication, in-house radix sort, tile ranges
// (upstream stages duplicateWithKeys / SortPairs / identifyTileRanges).
//
// The reference order is "stable ascending sort of key = (tile << 32) | depth_key over instances
// emitted in Gaussian-index order".  Instead of a 6-pass global LSD sort (144 B of HBM traffic per
// instance) the sort is done MSD-first, shaped to the key:
//   digit 0 = the whole tile field.  A counting sort on it needs only the per-tile histogram
//             (atomics in the preprocess kernel), one exclusive scan over T tiles — which IS
//             identifyTileRanges — and one scatter pass (8 B written per instance);
//   digit 1 = the 32-bit depth key, sorted per tile bucket inside shared memory.
// Bit-exactness: each Gaussian emits at most one instance per tile, so "emission order within a
// tile" is "ascending Gaussian id"; sorting each bucket by the 64-bit composite
// (depth_key << 32 | id) therefore reproduces the stable order exactly, whatever order the scatter
// atomics land in.
#include "common.cuh"

namespace gsvc {

// ---- exclusive scan over tiles: offsets, ranges, num_rendered; resets the scatter cursors ----------
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_kernel(int T, ImageView im)
{
    __shared__ unsigned long long warp_sums[SCAN_THREADS / 32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0ull;
    __syncthreads();
    // Coalesced chunks of SCAN_THREADS tiles; the running carry makes it a full scan for any T.
    for (int base = 0; base < T; base += SCAN_THREADS) {
        const int t = base + tid;
        const unsigned int c = t < T ? im.tile_count[t] : 0u;
        unsigned long long v = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += n;
        }
        if (lane == 31) warp_sums[wid] = v;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long n = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += n;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const unsigned long long incl = carry + (wid ? warp_sums[wid - 1] : 0ull) + v;
        if (t < T) {
            const unsigned long long excl = incl - c;
            im.tile_offset[t] = (unsigned int)excl;
            im.tile_cursor[t] = 0u;
            im.ranges[t] = c ? make_uint2((unsigned int)excl, (unsigned int)incl) : make_uint2(0u, 0u);
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) {
        im.hdr->num_rendered = carry_s;
        im.hdr->overflow = 0u;
    }
}

cudaError_t launch_tile_scan(const DevSettings& s, ImageView im, cudaStream_t st)
{
    tile_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(s.gx * s.gy, im);
    count_launch();
    return cudaGetLastError();
}

// ---- scatter: one instance (depth_key << 32 | id) per (Gaussian, tile in rect) into its tile bucket ----
__global__ void __launch_bounds__(256) scatter_kernel(int P, int gx, GeomView geo, ImageView im, BinView bin,
                                                      unsigned long long cap)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    const ushort4 r = geo.rect[g];
    if (r.z <= r.x || r.w <= r.y) return;
    const unsigned long long item = ((unsigned long long)ordered_u32(geo.feat2[g].y) << 32) | (unsigned int)g;
    for (int ty = r.y; ty < r.w; ty++)
        for (int tx = r.x; tx < r.z; tx++) {
            const int t = ty * gx + tx;
            const unsigned long long slot = (unsigned long long)im.tile_offset[t] + atomicAdd(im.tile_cursor + t, 1u);
            if (slot < cap)
                bin.inst[slot] = item;
            else
                im.hdr->overflow = 1u;
        }
}

cudaError_t launch_scatter(const DevSettings& s, int P, GeomView g, ImageView im, BinView b, long long cap,
                           cudaStream_t st)
{
    if (P <= 0) return cudaSuccess;
    scatter_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, s.gx, g, im, b, (unsigned long long)cap);
    count_launch();
    return cudaGetLastError();
}

// ---- per-tile sort of the 64-bit composites -----------------------------------------------------
// Normalised bitonic network: every compare-exchange moves the minimum to the lower index, so an
// arbitrary length n is handled by treating indices >= n as +inf (pairs that touch them are no-ops).
constexpr int SORT_THREADS = 128;
constexpr int SORT_SMEM_ITEMS = 4096;  // 32 KB; longer buckets are sorted in place in global memory

template <typename Ptr>
__device__ __forceinline__ void bitonic_sort(Ptr a, int n, int tid, int nthreads)
{
    int npad = 1;
    while (npad < n) npad <<= 1;
    const int half = npad >> 1;
    for (int k = 2; k <= npad; k <<= 1) {
        const int hk = k >> 1;
        for (int i = tid; i < half; i += nthreads) {
            const int blk = i / hk, pos = i - blk * hk;
            const int lo = blk * k + pos, hi = blk * k + k - 1 - pos;
            if (hi < n) {
                const unsigned long long x = a[lo], y = a[hi];
                if (x > y) { a[lo] = y; a[hi] = x; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j >= 1; j >>= 1) {
            for (int i = tid; i < half; i += nthreads) {
                const int lo = 2 * j * (i / j) + (i % j), hi = lo + j;
                if (hi < n) {
                    const unsigned long long x = a[lo], y = a[hi];
                    if (x > y) { a[lo] = y; a[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_tiles_kernel(ImageView im, BinView bin, unsigned long long cap)
{
    __shared__ unsigned long long s_items[SORT_SMEM_ITEMS];
    const int t = blockIdx.x;
    const uint2 rg = im.ranges[t];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0 || (unsigned long long)rg.y > cap) return;
    unsigned long long* src = bin.inst + rg.x;
    const int tid = threadIdx.x;
    if (n <= SORT_SMEM_ITEMS) {
        for (int i = tid; i < n; i += SORT_THREADS) s_items[i] = src[i];
        __syncthreads();
        bitonic_sort(s_items, n, tid, SORT_THREADS);
        for (int i = tid; i < n; i += SORT_THREADS) {
            const unsigned long long v = s_items[i];
            bin.point_list[rg.x + i] = (unsigned int)v;
            bin.depth_keys[rg.x + i] = (unsigned int)(v >> 32);
        }
    } else {
        __syncthreads();
        bitonic_sort(src, n, tid, SORT_THREADS);  // global-memory fallback (block-wide barriers order the accesses)
        for (int i = tid; i < n; i += SORT_THREADS) {
            const unsigned long long v = src[i];
            bin.point_list[rg.x + i] = (unsigned int)v;
            bin.depth_keys[rg.x + i] = (unsigned int)(v >> 32);
        }
    }
}

cudaError_t launch_sort_tiles(const DevSettings& s, ImageView im, BinView b, long long cap, cudaStream_t st)
{
    const int T = s.gx * s.gy;
    if (T <= 0) return cudaSuccess;
    sort_tiles_kernel<<<T, SORT_THREADS, 0, st>>>(im, b, (unsigned long long)cap);
    count_launch();
    return cudaGetLastError();
}

// ---- export for the bit-exact stage tests --------------------------------------------------------
__global__ void export_keys_kernel(int T, ImageView im, BinView bin, unsigned long long* sorted_keys,
                                   unsigned int* point_list, unsigned int* ranges)
{
    const int t = blockIdx.x;
    if (t >= T) return;
    const uint2 rg = im.ranges[t];
    if (ranges && threadIdx.x == 0) { ranges[2 * t] = rg.x; ranges[2 * t + 1] = rg.y; }
    for (unsigned int i = rg.x + threadIdx.x; i < rg.y; i += blockDim.x) {
        if (sorted_keys) sorted_keys[i] = ((unsigned long long)t << 32) | bin.depth_keys[i];
        if (point_list) point_list[i] = bin.point_list[i];
    }
}

cudaError_t launch_export_keys(const DevSettings& s, ImageView im, BinView b, long long R,
                               unsigned long long* sorted_keys, unsigned int* point_list, unsigned int* ranges,
                               cudaStream_t st)
{
    (void)R;
    const int T = s.gx * s.gy;
    export_keys_kernel<<<T, 128, 0, st>>>(T, im, b, sorted_keys, point_list, ranges);
    count_launch();
    return cudaGetLastError();
}

__global__ void export_geom_kernel(int P, GeomView geo, float* depth, float* xy, float* conic_opacity, float* rgb,
                                   int32_t* rect)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    const ushort4 r = geo.rect[g];
    const bool vis = r.z > r.x && r.w > r.y;
    float4 f0 = make_float4(0, 0, 0, 0), f1 = f0, f2 = f0;
    if (vis) { f0 = geo.feat0[g]; f1 = geo.feat1[g]; f2 = geo.feat2[g]; }
    if (depth) depth[g] = f2.y;
    if (xy) { xy[2 * g] = f0.x; xy[2 * g + 1] = f0.y; }
    if (conic_opacity) {
        conic_opacity[4 * g] = f0.z; conic_opacity[4 * g + 1] = f0.w;
        conic_opacity[4 * g + 2] = f1.x; conic_opacity[4 * g + 3] = f1.y;
    }
    if (rgb) { rgb[3 * g] = f1.z; rgb[3 * g + 1] = f1.w; rgb[3 * g + 2] = f2.x; }
    if (rect) { rect[4 * g] = r.x; rect[4 * g + 1] = r.y; rect[4 * g + 2] = r.z; rect[4 * g + 3] = r.w; }
}

cudaError_t launch_export_geom(int P, GeomView g, float* depth, float* xy, float* conic_opacity, float* rgb,
                               int32_t* rect, cudaStream_t st)
{
    if (P <= 0) return cudaSuccess;
    export_geom_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, g, depth, xy, conic_opacity, rgb, rect);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gsvc
