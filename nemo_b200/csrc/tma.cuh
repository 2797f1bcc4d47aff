// Asynchronous global -> shared copies of sm_90+/sm_100a (the TMA unit) and the mbarrier that tracks them, as inline PTX:
//   * bulk_g2s       cp.async.bulk             contiguous bytes (rows of a map / of a half spectrum), SASS UBLKCP
//   * tma_load_2d    cp.async.bulk.tensor.2d   a box of a 2-D tensor described by a CUtensorMap (column tiles), SASS UTMALDG
// One elected thread arms the barrier with the byte count (mbar_arrive_expect_tx) and issues the copies; every consumer waits on
// the barrier's phase parity.  The copied bytes arrive through the async proxy: before a buffer that threads have READ is handed
// back to the TMA unit, the readers synchronise and the issuing thread executes fence_proxy_async().
//
// Host side: make_tmap_2d() encodes a tensor map through the driver entry point (cudaGetDriverEntryPoint, so the library keeps
// linking against the static runtime only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barrier visible to the async proxy (the TMA unit completes transactions on it)
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// box of the tensor map at element coordinates (x = fastest dimension, y); out-of-range elements arrive as zeros and still count
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

// shared -> global box store (bulk async-group completion), SASS UTMASTG
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the stores issued so far have finished READING shared memory (the buffer may be rewritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tma

// 2-D tensor map over `rows` x `cols` elements of `elemBytes` (4 or 8: float / double) with a row pitch of `pitchBytes` (a
// multiple of 16; base 16-byte aligned), boxes of boxRows x boxCols elements, no swizzle, zero fill outside.  Returns 0 on success.
int nb_make_tmap_2d(CUtensorMap* map, const void* base, int elemBytes, unsigned long long rows, unsigned long long cols,
                    unsigned long long pitchBytes, unsigned boxRows, unsigned boxCols);
