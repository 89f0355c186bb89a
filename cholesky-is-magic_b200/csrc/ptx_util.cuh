// Thin inline-PTX wrappers used by the sm_100a kernels: mbarrier, TMA (cp.async.bulk[.tensor]),
// and the FP64 tensor-core instruction (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4).
//
// tcgen05.mma has no f64 kind on sm_100a (ptxas rejects .kind::f64), so the FP64 tensor path is
// mma.sync DMMA with register accumulators; operands are staged in shared memory by TMA.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>

namespace nes {

// cudaFuncSetAttribute is per DEVICE: one process may open contexts on several GPUs (nes_set_device +
// nes_start), so "configured" bits are kept per device ordinal and guarded against concurrent host threads.
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    // returns true when the caller must run the configuration for the current device (mutex held until
    // finish() is called)
    bool begin(int* dev_out) {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        mu.lock();
        if (done[dev]) {
            mu.unlock();
            return false;
        }
        *dev_out = dev;
        return true;
    }
    void finish(int dev, bool ok) {
        if (ok) done[dev] = true;
        mu.unlock();
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Release of a pipeline stage that the hardware cannot perform before the stage's fragment loads have RETURNED.
//
// Why: the consumers read a TMA-filled stage with ordinary loads, run DMMAs on the fragments and then arrive on the
// stage's "empty" barrier, after which the producer lets TMA overwrite the stage.  dmma884 is a non-volatile asm,
// so the compiler sinks the last DMMAs below the arrive: in SASS the arrive then follows the last fragment LOAD
// ISSUE, not its completion (SYNCS.ARRIVE takes no register the loads produce, so no scoreboard holds it back).
// Normally a load returns long before a TMA refill can land; but with a second CTA on the SM (or the same CTA's
// epilogue) keeping the load/store unit busy with global read-modify-write traffic a fragment load can be late
// by more than the refill latency and then reads the NEXT occupant of the stage: one 8-row fragment of one
// tile wrong, about once per 10^5 tiles (found in round 2 as "lost updates" of the half-tile update kernel; the
// evidence trail is in DESIGN.md section 5).  Cure: a true data dependency.  `dep` must be derived from
// every fragment register loaded from the stage; `never` is a run-time value that is always 0 but
// that neither nvcc nor ptxas can prove to be 0, so `dep & never` costs two integer instructions, keeps the
// dependency alive and leaves the address unchanged.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t dep, uint32_t never) {
    asm volatile(
        "{\n"
        ".reg .b32 a;\n"
        "and.b32 a, %1, %2;\n"
        "add.u32 a, a, %0;\n"
        "mbarrier.arrive.shared::cta.b64 _, [a];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(dep), "r"(never)
        : "memory");
}

// Dependency value: XOR of the high words of every fragment register loaded from the stage (8 row + 4 column
// fragments per k-step).  The arrive then waits for the LOADS only -- issued long before, so normally complete --
// and not for the DMMA pipeline to drain (depending on the accumulators instead cost 3.4% of the formation:
// every warp stalled for the latency of its last DMMA at each chunk boundary).
__device__ __forceinline__ uint32_t frag_dependency(uint32_t d, const double (&af)[8], const double (&bf)[4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d ^= static_cast<uint32_t>(__double2hiint(af[i]));
#pragma unroll
    for (int j = 0; j < 4; ++j) d ^= static_cast<uint32_t>(__double2hiint(bf[j]));
    return d;
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 2D tiled TMA load: box -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%2, %3}], [%4];"
        ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// 2D tiled TMA store: shared -> global (bulk async group).  Writes by ordinary threads must be
// published to the async proxy first: fence_proxy_async() by the writers, then a barrier.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(src))
                 : "memory");
}

__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// 1D bulk copy global -> shared (bytes multiple of 16, both sides 16B aligned).
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 1D bulk copy shared -> global (bulk async group; bytes multiple of 16, both sides 16B aligned).
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst)),
                 "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// A tensor map that lives in GLOBAL memory (written by a copy, i.e. through the generic proxy) must be
// acquired by the tensormap proxy before cp.async.bulk.tensor may read it.
__device__ __forceinline__ void fence_tensormap_acquire(const CUtensorMap* map) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(map))
                 : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// D(8x8) += A(8x4, row) * B(4x8, col).  Lane l: g = l>>2, t = l&3.
//   a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace nes
