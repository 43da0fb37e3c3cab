// tcgen05 / TMEM primitives (sm_100a inline PTX) and the operand layout of the large-minibatch tensor-core tiles.
//
// Arithmetic: fp32 operands are split a = hi + lo with hi = RN_tf32(a), lo = RN_tf32(a - hi), and a product is formed
// as hi*hi + hi*lo + lo*hi on the tensor cores (kind::tf32, fp32 accumulate in TMEM): "3xTF32", ~2^-22 relative per
// product, which keeps the 1e-4 loss contract of north_star (plain TF32 is ~2^-11 and does not).
//
// Shared-memory operand layouts (E = MN extent, Kx = K extent of the image, floats):
//   K-major  (SWIZZLE_NONE / "interleave"): 16-byte granules of 4 consecutive k; element (e, k) lives in granule
//            (k/4)*E + e.  Eight consecutive granules (8 rows x 16 bytes) are one UMMA core matrix.
//            descriptor: LBO = 16*E (next 4 k), SBO = 128 (next 8 rows); one kind::tf32 MMA (K = 8) advances 2*16*E bytes.
//   MN-major (SWIZZLE_128B_BASE32B -- the ONLY layout the tensor core accepts for MN-major 32-bit operands; with any
//            other layout type the MMA silently produces zeros.  Address function decoded on B200 with
//            profiles/mma_probe.py): atoms of 4 k-rows x 32 consecutive e (128 bytes per row, 512 bytes per atom); row
//            k%4 of an atom stores its four 32-byte chunks (8 floats) at chunk position c ^ (k%4), i.e. byte-address bits
//            [5,7) ^= bits [7,9).  Atoms are laid out [k/4][e/32]:
//            float offset = ((k/4)*(E/32) + e/32)*128 + (k%4)*32 + ((e%32) ^ ((k%4) << 3))
//            descriptor: LBO = 512 (next 32 e), SBO = 16*E (next 4 k), layout type 1; one MMA (K = 8) advances 2*SBO bytes.
//            An aligned quad of 4 consecutive e is one 16-byte granule in natural order.
//            A K-chunk of an image is contiguous in both layouts; images must be 1024-byte aligned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg {
namespace mma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- TMEM allocation (one full warp executes these) ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- descriptors --------------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100): bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) = 1, [61,64) layout type (0 = no swizzle, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
    uint64_t d = (uint64_t)(layout_type & 7u) << 61;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::tf32, fp32 accumulate: [4,6) c=F32(1), [7,10) a=TF32(2), [10,13) b=TF32(2),
// 15 a MN-major, 16 b MN-major, [17,23) N>>3, [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T   (one thread issues)
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// ---- TMEM -> registers: warp w reads lanes 32*(w%4) .. +31, thread i gets lane i of that quarter, 16 consecutive columns --
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// float offset of element (e, k) in the images described at the top of this file
__host__ __device__ __forceinline__ int kmajor_off(int e, int k, int E) { return (((k >> 2) * E + e) << 2) + (k & 3); }
__host__ __device__ __forceinline__ int mnmajor_off(int e, int k, int E) {
    return (((k >> 2) * (E >> 5) + (e >> 5)) << 7) + ((k & 3) << 5) + ((e & 31) ^ ((k & 3) << 3));
}

// ---- 3xTF32 operand split ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}
__device__ __forceinline__ void split4(float4 x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

}  // namespace mma
}  // namespace sg
