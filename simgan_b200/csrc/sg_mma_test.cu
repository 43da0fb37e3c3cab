// Self-test of the tcgen05 building blocks (no reference counterpart): one CTA forms D(M,N) = A . B^T on the tensor
// cores from operands it lays out itself in the granule layout of sg_mma.cuh, in every operand-major combination the
// large-minibatch tiles use.  tests/test_gpu_mma.py compares it with a float64 product.
#include "sg_common.cuh"
#include "sg_mma.cuh"

namespace sg {

struct MmaTestArgs {
    int M, N, K;
    int a_mn, b_mn;       // operand given MN-major ((K, M) / (K, N) row-major) instead of K-major ((M, K) / (N, K))
    int a_direct;         // A (K-major) is kept as an UNSPLIT fp32 master [k/4][row] with a row pitch of M+1 granules; the hi
                          // pass reads it in place (the tensor core drops the 13 low mantissa bits), only lo is an image
    int raw;              // probe: A / B are verbatim shared-memory images and the descriptor strides come from rs[]
    int rs[8];            // {a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step} in bytes, {a_layout_type, b_layout_type}
    int passes;           // 3 = 3xTF32 (hi*hi + hi*lo + lo*hi), 1 = plain TF32
    const float *A, *B;
    float* D;
};

// operand (E x Kx) -> hi / lo granule buffers
__device__ void fill_operand(float4* hi, float4* lo, const float* __restrict__ src, int E, int Kx, int mn_major, int tid, int nth) {
    if (!mn_major) {
        // source (e, k) row-major; granule (k/4)*E + e
        const int KB = Kx >> 2;
        for (int g = tid; g < KB * E; g += nth) {
            const int kb = g / E, e = g - kb * E;
            const float4 x = *reinterpret_cast<const float4*>(src + (size_t)e * Kx + 4 * kb);
            float4 h, l;
            mma::split4(x, h, l);
            hi[g] = h; lo[g] = l;
        }
    } else {
        // source (k, e) row-major -> SWIZZLE_128B_BASE32B image: the aligned quad e..e+3 is one granule
        const int EB = E >> 2;
        for (int q = tid; q < EB * Kx; q += nth) {
            const int k = q / EB, e = 4 * (q - k * EB);
            const float4 x = *reinterpret_cast<const float4*>(src + (size_t)k * E + e);
            float4 h, l;
            mma::split4(x, h, l);
            const int off = mma::mnmajor_off(e, k, E);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(hi) + off) = h;
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(lo) + off) = l;
        }
    }
}

__global__ void __launch_bounds__(128, 1) mma_selftest_kernel(MmaTestArgs a) {
    extern __shared__ __align__(1024) float4 smem4[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = a.M, N = a.N, K = a.K;
    float4* Ahi = smem4;
    const int apitch = a.a_direct ? M + 1 : M;                      // granules between consecutive k-quads of A's hi operand
    float4* Alo = Ahi + round_up((K / 4) * apitch, 64);             // images stay 1024-byte aligned
    float4* Bhi = Alo + (size_t)M * K / 4;
    float4* Blo = Bhi + (size_t)N * K / 4;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;

    if (warp == 0) {
        mma::tmem_alloc(&tmem_base_s, ncols);
        mma::tmem_relinquish();
    }
    if (tid == 0) mbar_init(&bar, 1);
    if (a.raw) {
        for (int g = tid; g < M * K / 4; g += 128) Ahi[g] = reinterpret_cast<const float4*>(a.A)[g];
        for (int g = tid; g < N * K / 4; g += 128) Bhi[g] = reinterpret_cast<const float4*>(a.B)[g];
    } else {
        if (a.a_direct) {
            for (int g = tid; g < (K / 4) * M; g += 128) {
                const int kb = g / M, e = g - kb * M;
                const float4 x = *reinterpret_cast<const float4*>(a.A + (size_t)e * K + 4 * kb);
                Ahi[kb * apitch + e] = x;
                float4 h;
                h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                Alo[g] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
            }
        } else {
            fill_operand(Ahi, Alo, a.A, M, K, a.a_mn, tid, 128);
        }
        fill_operand(Bhi, Blo, a.B, N, K, a.b_mn, tid, 128);
    }
    mma::fence_async_smem();
    mma::fence_before_sync();
    __syncthreads();
    mma::fence_after_sync();
    const uint32_t tbase = tmem_base_s;

    if (warp == 0) {
        if (mma::elect_one()) {
            const uint32_t idesc = mma::make_idesc_tf32(M, N, a.a_mn, a.b_mn);
            uint32_t a_lbo = a.a_mn ? 512u : 16u * M, a_sbo = a.a_mn ? 16u * M : 128u, a_step = 32u * M;
            uint32_t b_lbo = a.b_mn ? 512u : 16u * N, b_sbo = a.b_mn ? 16u * N : 128u, b_step = 32u * N;
            uint32_t a_lt = a.a_mn ? 1u : 0u, b_lt = a.b_mn ? 1u : 0u;
            if (a.raw) {
                a_lbo = a.rs[0]; a_sbo = a.rs[1]; a_step = a.rs[2]; b_lbo = a.rs[3]; b_sbo = a.rs[4]; b_step = a.rs[5];
                a_lt = a.rs[6]; b_lt = a.rs[7];
            }
            const uint32_t ah = mma::smem_addr(Ahi), al = mma::smem_addr(Alo), bh = mma::smem_addr(Bhi), bl = mma::smem_addr(Blo);
            const uint32_t ah_lbo = a.a_direct ? 16u * apitch : a_lbo, ah_step = a.a_direct ? 32u * apitch : a_step;
            uint32_t accum = 0;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t dAh = mma::make_desc(ah + ks * ah_step, ah_lbo, a_sbo, a_lt), dAl = mma::make_desc(al + ks * a_step, a_lbo, a_sbo, a_lt);
                const uint64_t dBh = mma::make_desc(bh + ks * b_step, b_lbo, b_sbo, b_lt), dBl = mma::make_desc(bl + ks * b_step, b_lbo, b_sbo, b_lt);
                if (a.passes == 3) {
                    mma::mma_tf32(tbase, dAl, dBh, idesc, accum); accum = 1;
                    mma::mma_tf32(tbase, dAh, dBl, idesc, accum);
                }
                mma::mma_tf32(tbase, dAh, dBh, idesc, accum); accum = 1;
            }
            mma::commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    mma::fence_after_sync();

    // epilogue: M = 128: row m in lane m; M = 64: row m in lane (m % 16) + 32 * (m / 16)
    int row;
    bool live;
    if (M == 128) { row = 32 * warp + lane; live = true; }
    else { row = 16 * warp + lane; live = lane < 16; }
    for (int c = 0; c < N; c += 16) {
        float v[16];
        mma::tmem_ld16(tbase + ((uint32_t)(32 * warp) << 16) + (uint32_t)c, v);
        mma::tmem_ld_wait();
        if (live) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c + j < N) a.D[(size_t)row * N + c + j] = v[j];
        }
    }
    mma::fence_before_sync();
    __syncthreads();
    if (warp == 0) mma::tmem_dealloc(tbase, ncols);
}

}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

int sg_selftest_mma(int M, int N, int K, int a_mn_major, int b_mn_major, int passes, const float* A, const float* B, float* D,
                    const int* h_raw_strides, void* stream) {
    SG_REQUIRE(M == 64 || M == 128, "sg_selftest_mma: M must be 64 or 128");
    SG_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "sg_selftest_mma: N must be a multiple of 16 in [16,256]");
    SG_REQUIRE(K >= 8 && K % 8 == 0, "sg_selftest_mma: K must be a positive multiple of 8");
    SG_REQUIRE(passes == 1 || passes == 3, "sg_selftest_mma: passes must be 1 or 3");
    SG_REQUIRE(A && B && D, "sg_selftest_mma: null pointer");
    const bool a_direct = a_mn_major == 2;
    if (a_direct) a_mn_major = 0;
    SG_REQUIRE(!a_direct || (!h_raw_strides && passes == 3), "sg_selftest_mma: the in-place hi operand is a 3-pass, non-raw variant");
    const size_t smem = (size_t)2 * (M + N) * K * sizeof(float) + (a_direct ? (size_t)(K / 4 + 64) * 16 : 0);
    SG_REQUIRE(smem <= 200 * 1024, "sg_selftest_mma: operands need %zu bytes of shared memory", smem);
    SG_REQUIRE(!h_raw_strides || passes == 1, "sg_selftest_mma: raw images run a single pass");
    SG_REQUIRE(h_raw_strides || ((!a_mn_major || M % 32 == 0) && (!b_mn_major || N % 32 == 0)),
               "sg_selftest_mma: an MN-major operand needs an MN extent that is a multiple of 32");
    MmaTestArgs a{M, N, K, a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, a_direct ? 1 : 0, h_raw_strides ? 1 : 0, {0, 0, 0, 0, 0, 0, 0, 0}, passes, A, B, D};
    if (h_raw_strides)
        for (int i = 0; i < 8; ++i) a.rs[i] = h_raw_strides[i];
    SG_CUDA(cudaFuncSetAttribute(mma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

#pragma GCC visibility pop
}
