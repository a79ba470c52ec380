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
#include "common.cuh"

namespace gsvc {

constexpr int AR_THREADS = 512;

// Handshake of this rank with every peer.  slot(owner, writer) is one word of `owner`'s pad that only `writer` raises
// and only `owner` lowers: raise = wait until it is 0, set it to 1; lower = wait until it is 1, set
// it back to 0.  Stateless (the pad is all zeros between two launches), so a CUDA graph can replay it.
__device__ __forceinline__ unsigned int cas_release_sys(unsigned int* a, unsigned int cmp, unsigned int val)
{
    unsigned int old;
    asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ unsigned int cas_acquire_sys(unsigned int* a, unsigned int cmp, unsigned int val)
{
    unsigned int old;
    asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}

__device__ __forceinline__ void peer_handshake(unsigned int* const* pads, int rank, int world)
{
    if ((int)threadIdx.x < world) {
        const int peer = (int)threadIdx.x;
        unsigned int* theirs = pads[peer] + rank;      // I raise it, the peer lowers it
        unsigned int* mine = pads[rank] + peer;        // the peer raises it, I lower it
        while (cas_release_sys(theirs, 0u, 1u) != 0u) {}
        while (cas_acquire_sys(mine, 1u, 0u) != 1u) {}
    }
}

// ONE CTA of the launch talks to the peers (world words of pad traffic per rank and handshake, whatever the grid size;
// with a handshake per CTA the 8-GPU exchange lost 15 us between 16 and 128 CTAs).  state = {go, done}: two words of
// LOCAL device memory, zero between launches.
//   begin: CTA 0 handshakes (every peer's producer kernels have finished), then opens `go` for the other CTAs.
__device__ __forceinline__ void exchange_begin(unsigned int* const* pads, unsigned int* state, int rank, int world)
{
    if (blockIdx.x == 0) {
        peer_handshake(pads, rank, world);
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(state), "r"(1u) : "memory");
    } else {
        if (threadIdx.x == 0) {
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(state) : "memory");
            } while (v != 1u);
        }
        __syncthreads();
    }
}
//   end: every CTA's stores are performed system-wide, then the LAST CTA to get here resets the state and handshakes
//   (every peer's stores into this rank's buffer have landed before the kernel completes).
__device__ __forceinline__ void exchange_end(unsigned int* const* pads, unsigned int* state, int rank, int world)
{
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(prev) : "l"(state + 1), "r"(1u) : "memory");
        s_last = prev == gridDim.x - 1;
        if (s_last) {
            state[0] = 0u;
            state[1] = 0u;
            __threadfence();
        }
    }
    __syncthreads();
    if (s_last) peer_handshake(pads, rank, world);
}

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
            if (i < hi4)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < hi4)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(mc + i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
        }
    }
    exchange_end(pads, state, rank, world);
}

// The same exchange without a multicast mapping (or for two ranks, where sending one's own copy through the switch
// costs more than it saves): rank r reads its slice from every rank's buffer with peer loads (NVLink), adds the copies
// in rank order — one adder per element, so every rank receives the SAME sum — and stores the result into every
// rank's buffer with peer stores.
__device__ __forceinline__ float4 ld_sys(const float4* p)
{
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(float4* p, float4 v)
{
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

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
    exchange_end(pads, state, rank, WORLD);
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
