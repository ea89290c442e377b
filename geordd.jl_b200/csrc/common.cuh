// geordd_b200 -- shared device helpers (sm_100a only).
//
// FP64 on Blackwell reaches the tensor pipe through warp-level mma.sync (SASS DMMA.8x8x4);
// tcgen05 has no f64 kind.  Everything dense in this library is built from the
// m8n8k4 f64 MMA below, fed from shared memory staged by cp.async (LDGSTS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.h"

namespace geordd {

constexpr int NTHREADS = 256;    // 8 warps per CTA
constexpr int PAD = 4;           // smem row padding (doubles): stride == 4 (mod 16) -> conflict-free LDS.64 fragments

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col) ---------------------------------
// lane l: g = l>>2, t = l&3.  a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- cp.async (16 B, L2 only: tiles are produced by other CTAs, never trust L1) ----------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- dataflow flags ------------------------------------------------------------------------------
// A tile is published with st.release.gpu after __threadfence()+__syncthreads(); consumers poll with
// ld.acquire.gpu.  Flags carry the launch epoch so they never need clearing.
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_int(const int* p) {   // generic address (global or shared)
    return *reinterpret_cast<const volatile int*>(p);
}

// Spin (one thread) until *flag == epoch.  A watchdog turns a scheduling bug into an error code
// instead of a hung GPU: after ~4 s (at ~2 GHz) the abort word is set and every waiter returns false.
constexpr long long WATCHDOG_CYCLES = 8000000000LL;
__device__ __forceinline__ bool wait_flag(const unsigned* flag, unsigned epoch, int* abort_word) {
    if (ld_acquire(flag) == epoch) return true;
    long long t0 = clock64();
    unsigned ns = 20;
    while (true) {
        if (ld_acquire(flag) == epoch) return true;
        if (ld_volatile_int(abort_word) != 0) return false;
        if (clock64() - t0 > WATCHDOG_CYCLES) {
            atomicCAS(abort_word, 0, -9);
            return false;
        }
        __nanosleep(ns);
        if (ns < 200) ns += 20;
    }
}

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }


// ---- mbarrier + bulk async copy (TMA unit, SASS UBLKCP) ---------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// Orders generic-proxy accesses (ld/st, flag acquire) before subsequent async-proxy (bulk copy) accesses.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive whose address depends on `dep_zero` (always 0 at run time): the instruction cannot issue before the
// registers that fed dep_zero are written, i.e. before the shared-memory loads that produced them have returned.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, unsigned dep_zero) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar) + dep_zero) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Raise the expected transaction bytes of the current phase without arriving.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// Wait for the phase with the given parity.  Polls the abort word so that an error / watchdog anywhere
// releases every waiter (returns false).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity, int* abort_word) {
    if (mbar_try_wait(bar, parity)) return true;
    long long t0 = clock64();
    unsigned spins = 0;
    while (true) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((++spins & 15u) == 0) {
            if (ld_volatile_int(abort_word) != 0) return false;
            if (clock64() - t0 > WATCHDOG_CYCLES) {
                atomicCAS(abort_word, 0, -9);
                return false;
            }
        }
    }
}
// Make the completion of all cp.async (LDGSTS) issued so far by this thread count as one (pre-counted) arrival.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Contiguous global -> shared copy by the TMA unit; completion is signalled on `bar` (complete_tx).
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace geordd
