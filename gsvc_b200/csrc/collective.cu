// Sum all-reduce of the frame-sharded step's gradient buffer through the NVSwitch
// (the exchange step of SURVEY.md §8e: every rank rasterizes its own frames of the TSW window, the [P,14] parameter
//  gradients are summed over the ranks — what DistributedDataParallel would do for pipeline/train.py:462).
//
// The buffer of every rank lives in symmetric memory that is also mapped as ONE multicast object.  A load-reduce
// through the multicast address makes the switch fetch the addressed 16 bytes from every GPU and return their sum; a
// store through it writes all GPUs.  Rank r therefore owns the r-th slice of the buffer: it reads the slice's SUM
// with multimem.ld_reduce and broadcasts it with multimem.st — each GPU's links carry the buffer once out and once
// in, no GPU receives N copies, and no SM adds anything.  Everything is ONE launch: the ranks handshake before (the
// peers' producer kernels have finished) and after (their stores have landed) through one word per pair of ranks in
// the symmetric signal pad.
#include "collective.cuh"

namespace gsvc {

template <int AR_UNROLL>
__global__ void __launch_bounds__(AR_THREADS)
switch_allreduce_kernel(float4* __restrict__ mc, unsigned int* const* __restrict__ pads, unsigned int* state, int rank,
                        int world, long long lo4, long long hi4)
{
    exchange_begin(pads, state, rank, world);
    const long long stride = (long long)gridDim.x * AR_THREADS;
    for (long long i0 = lo4 + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi4; i0 += stride * AR_UNROLL) {
        float4 v[AR_UNROLL];
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < hi4) v[u] = mc_ld_reduce(mc + i);
        }
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < hi4) mc_st(mc + i, v[u]);
        }
    }
    exchange_end(pads, state, rank, world, gridDim.x);
}

// The same exchange without a multicast mapping (or for two ranks, where sending one's own copy through the switch
// costs more than it saves): rank r reads its slice from every rank's buffer with peer loads (NVLink), adds the copies
// in rank order — one adder per element, so every rank receives the SAME sum — and stores the result into every
// rank's buffer with peer stores.
template <int WORLD, int UNROLL>
__global__ void __launch_bounds__(AR_THREADS)
peer_allreduce_kernel(float4* const* __restrict__ bufs, unsigned int* const* __restrict__ pads, unsigned int* state, int rank,
                      long long lo4, long long hi4)
{
    exchange_begin(pads, state, rank, WORLD);
    float4* b[WORLD];
#pragma unroll
    for (int q = 0; q < WORLD; q++) b[q] = bufs[q];
    const long long stride = (long long)gridDim.x * AR_THREADS;
    for (long long i0 = lo4 + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi4; i0 += stride * UNROLL) {
        float4 v[UNROLL][WORLD];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < hi4) {
#pragma unroll
                for (int q = 0; q < WORLD; q++) v[u][q] = ld_sys(b[q] + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < hi4) {
                float4 a = v[u][0];
#pragma unroll
                for (int q = 1; q < WORLD; q++) { a.x += v[u][q].x; a.y += v[u][q].y; a.z += v[u][q].z; a.w += v[u][q].w; }
#pragma unroll
                for (int q = 0; q < WORLD; q++) st_sys(b[q] + i, a);
            }
        }
    }
    exchange_end(pads, state, rank, WORLD, gridDim.x);
}

cudaError_t launch_switch_allreduce(float* multicast, void* const* buffers, unsigned int* const* pads, unsigned int* state,
                                    int rank, int world, long long numel, int n_ctas, cudaStream_t st)
{
    const long long n4 = numel / 4;
    const long long lo4 = n4 * rank / world, hi4 = n4 * (rank + 1) / world;
    count_launch();
    if (multicast) {
        // (4 vector loads in flight per thread; 8 or 2 measured within 2 % of it at 28 MB)
        switch_allreduce_kernel<4><<<n_ctas, AR_THREADS, 0, st>>>(reinterpret_cast<float4*>(multicast), pads, state, rank,
                                                                  world, lo4, hi4);
        return cudaGetLastError();
    }
    float4* const* bufs = reinterpret_cast<float4* const*>(buffers);
    switch (world) {
    case 1: return cudaSuccess;
    case 2: peer_allreduce_kernel<2, 4><<<n_ctas, AR_THREADS, 0, st>>>(bufs, pads, state, rank, lo4, hi4); break;
    case 4: peer_allreduce_kernel<4, 2><<<n_ctas, AR_THREADS, 0, st>>>(bufs, pads, state, rank, lo4, hi4); break;
    case 8: peer_allreduce_kernel<8, 1><<<n_ctas, AR_THREADS, 0, st>>>(bufs, pads, state, rank, lo4, hi4); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace gsvc
