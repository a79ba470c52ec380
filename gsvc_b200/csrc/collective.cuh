// Device-side pieces of the rank-to-rank exchange (csrc/collective.cu has the stand-alone all-reduce, csrc/preprocess_bwd.cu
// the per-Gaussian backward that carries the exchange in the same launch).
#pragma once
#include "common.cuh"

namespace gsvc {

constexpr int AR_THREADS = 512;   // threads per CTA of the stand-alone all-reduce

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

// Every wait of the exchange is bounded: a rank that never arrives (a crashed peer process, a caller that skipped the
// collective call) must not leave this GPU spinning for good.  After ~20 s (clock64 runs at the SM clock) the wait gives
// up, counts itself in the sticky word state[EXCHANGE_TIMEOUT_WORD] and the kernel runs to its end — the buffer is then NOT
// the sum over the ranks; the owner reads that word after a synchronisation (SwitchAllReduce.timeouts()).
constexpr long long EXCHANGE_WAIT_CYCLES = 40000000000ll;
constexpr int EXCHANGE_TIMEOUT_WORD = 2 + GSVC_RAST_EXCHANGE_MAX_CHUNKS;   // state = go, done, chunk counters, this word

__device__ __forceinline__ long long sm_clock()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

struct SpinGuard {
    long long t0;
    unsigned int n;
    unsigned int* sticky;
    __device__ __forceinline__ explicit SpinGuard(unsigned int* state) : t0(sm_clock()), n(0u), sticky(state + EXCHANGE_TIMEOUT_WORD) {}
    // true: keep waiting
    __device__ __forceinline__ bool keep_waiting()
    {
        if ((++n & 0xfffu) != 0u) return true;                     // look at the clock every 4096 polls
        if (sm_clock() - t0 < EXCHANGE_WAIT_CYCLES) return true;
        atomicAdd(sticky, 1u);
        return false;
    }
};

__device__ __forceinline__ void peer_handshake(unsigned int* const* pads, unsigned int* state, int rank, int world)
{
    if ((int)threadIdx.x < world) {
        const int peer = (int)threadIdx.x;
        unsigned int* theirs = pads[peer] + rank;      // I raise it, the peer lowers it
        unsigned int* mine = pads[rank] + peer;        // the peer raises it, I lower it
        SpinGuard guard(state);
        while (cas_release_sys(theirs, 0u, 1u) != 0u && guard.keep_waiting()) {}
        while (cas_acquire_sys(mine, 1u, 0u) != 1u && guard.keep_waiting()) {}
    }
}

// ONE CTA of the launch talks to the peers (world words of pad traffic per rank and handshake, whatever the grid size;
// with a handshake per CTA the 8-GPU exchange lost 15 us between 16 and 128 CTAs).  state = {go, done}: two words of
// LOCAL device memory, zero between launches.
//   begin: CTA 0 handshakes (every peer's producer kernels have finished), then opens `go` for the other CTAs.
__device__ __forceinline__ void exchange_begin(unsigned int* const* pads, unsigned int* state, int rank, int world)
{
    if (blockIdx.x == 0) {
        peer_handshake(pads, state, rank, world);
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(state), "r"(1u) : "memory");
    } else {
        if (threadIdx.x == 0) {
            unsigned int v;
            SpinGuard guard(state);
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(state) : "memory");
            } while (v != 1u && guard.keep_waiting());
        }
        __syncthreads();
    }
}
//   end: every CTA's stores are performed system-wide, then the LAST CTA to get here resets the state and handshakes
//   (every peer's stores into this rank's buffer have landed before the kernel completes).
__device__ __forceinline__ void exchange_end(unsigned int* const* pads, unsigned int* state, int rank, int world,
                                             unsigned int n_ctas, int n_extra = 0)
{
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(prev) : "l"(state + 1), "r"(1u) : "memory");
        s_last = prev == n_ctas - 1;
        if (s_last) {
            for (int i = 0; i < 2 + n_extra; i++) state[i] = 0u;      // (go, done, and the caller's extra words)
            __threadfence();
        }
    }
    __syncthreads();
    if (s_last) peer_handshake(pads, state, rank, world);
}

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


// 16 bytes of the slice this rank owns, summed over the ranks and written back to all of them: through the switch
// (multicast mapping) or, without one, with peer loads (added in rank order: one adder per element, so every rank
// receives the same bits) and peer stores.
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- the exchange carried by ANOTHER kernel's launch (the per-Gaussian backward, csrc/preprocess_bwd.cu) ------------
// The first n_ex CTAs of that launch take this role, the others compute rows of the buffer, `chunk_ctas` CTAs per chunk,
// and count themselves into state[2 + chunk] when their rows are stored.  Chunk by chunk the exchange waits for the
// local count, handshakes with the peers (THEIR chunk is stored as well) and moves the chunk — so the transfer of the
// first rows runs under the computation of the last ones, and the launch ends when the summed buffer is in place.
constexpr int EXCHANGE_MAX_PEER_WORLD = 8;   // ranks of the peer-load path (their buffer addresses live in registers)

constexpr int MOVE_UNROLL = 8;   // 16-byte words in flight per mover thread (64 movers x 128 threads: 1 MB)

__device__ __forceinline__ void move_slice(const ExchangeArgs& ex, float4* const (&b)[EXCHANGE_MAX_PEER_WORLD],
                                           long long lo4, long long hi4, int cta)
{
    const long long stride = (long long)ex.n_ex * blockDim.x;
    for (long long i0 = lo4 + (long long)cta * blockDim.x + threadIdx.x; i0 < hi4; i0 += MOVE_UNROLL * stride) {
        float4 v[MOVE_UNROLL];
        if (ex.mc) {
#pragma unroll
            for (int u = 0; u < MOVE_UNROLL; u++)
                if (i0 + u * stride < hi4) v[u] = mc_ld_reduce(ex.mc + i0 + u * stride);
#pragma unroll
            for (int u = 0; u < MOVE_UNROLL; u++)
                if (i0 + u * stride < hi4) mc_st(ex.mc + i0 + u * stride, v[u]);
        } else {
#pragma unroll
            for (int u = 0; u < MOVE_UNROLL; u++) {
                const long long i = i0 + u * stride;
                if (i < hi4) {
                    float4 a = ld_sys(b[0] + i);
#pragma unroll
                    for (int q = 1; q < EXCHANGE_MAX_PEER_WORLD; q++) {
                        if (q < ex.world) {
                            const float4 c = ld_sys(b[q] + i);
                            a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
                        }
                    }
                    v[u] = a;
                }
            }
#pragma unroll
            for (int u = 0; u < MOVE_UNROLL; u++) {
                const long long i = i0 + u * stride;
                if (i < hi4) {
#pragma unroll
                    for (int q = 0; q < EXCHANGE_MAX_PEER_WORLD; q++)
                        if (q < ex.world) st_sys(b[q] + i, v[u]);
                }
            }
        }
    }
}

// Pad layout: words [0, world) are the handshake words (peer_handshake), word world + k * world + q says "rank q has
// stored its rows of chunk k".  CTA 0 of the role only coordinates: when the local count of chunk k is complete it
// raises this rank's word for chunk k in EVERY rank's pad (one store per peer, nothing to wait for) and goes on to the
// next chunk; the other CTAs poll the words of chunk k in their OWN pad (local memory) and move their share of the
// chunk as soon as all ranks have raised theirs.  The words are lowered again by the last CTA before the closing
// handshake, so the pad is all zeros between launches.
__device__ __forceinline__ void exchange_role(const ExchangeArgs& ex)
{
    const int W = ex.world;
    unsigned int* const my_flags = ex.pads[ex.rank] + W;
    if (blockIdx.x == 0) {
        for (int k = 0; k < ex.n_chunks; k++) {
            if (threadIdx.x == 0) {
                const unsigned int want = (unsigned int)min(ex.chunk_ctas, ex.n_compute - k * ex.chunk_ctas);
                unsigned int v;
                SpinGuard guard(ex.state);
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ex.state + 2 + k) : "memory");
                } while (v != want && guard.keep_waiting());
                __threadfence_system();
            }
            __syncthreads();
            if ((int)threadIdx.x < W)
                asm volatile("st.release.sys.global.u32 [%0], %1;"
                             ::"l"(ex.pads[threadIdx.x] + W + (size_t)k * W + ex.rank), "r"(1u) : "memory");
        }
    } else {
        ExchangeArgs mv = ex;
        mv.n_ex = ex.n_ex - 1;                      // CTAs that move data: 1 .. n_ex - 1
        float4* b[EXCHANGE_MAX_PEER_WORLD];         // (fetched once: between the volatile accesses of the loop every
#pragma unroll                                      //  use of ex.bufs[q] would be a dependent load of its own)
        for (int q = 0; q < EXCHANGE_MAX_PEER_WORLD; q++) b[q] = (!ex.mc && q < W) ? ex.bufs[q] : nullptr;
        for (int k = 0; k < ex.n_chunks; k++) {
            if ((int)threadIdx.x < W) {
                unsigned int v;
                SpinGuard guard(ex.state);
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + (size_t)k * W + threadIdx.x) : "memory");
                } while (v != 1u && guard.keep_waiting());
            }
            __syncthreads();
            const long long c_lo = (long long)k * ex.chunk_f4;
            const long long c_n = min(ex.chunk_f4, ex.total_f4 - c_lo);
            move_slice(mv, b, c_lo + c_n * ex.rank / W, c_lo + c_n * (ex.rank + 1) / W, (int)blockIdx.x - 1);
        }
    }
    // closing: every CTA of the role counts in; the last one lowers this rank's flag words and the local counters, then
    // handshakes (every peer's stores into this rank's buffer have landed when the launch completes)
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(prev) : "l"(ex.state + 1), "r"(1u) : "memory");
        s_last = prev == (unsigned int)ex.n_ex - 1;
    }
    __syncthreads();
    if (s_last) {
        for (int i = threadIdx.x; i < ex.n_chunks * W; i += blockDim.x) my_flags[i] = 0u;
        for (int i = threadIdx.x; i < 2 + ex.n_chunks; i += blockDim.x) ex.state[i] = 0u;
        __threadfence_system();
        __syncthreads();
        peer_handshake(ex.pads, ex.state, ex.rank, W);
    }
}

}  // namespace gsvc
