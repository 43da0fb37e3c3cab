// Large-minibatch PPO tile on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory).
//
// Replaces, for minibatches of thousands of rows, the CUDA-core tile phase of the persistent PPO kernel (sg_ppo.cu):
// the six contractions of A2C/model.py:255-264 + A2C/distributions.py:110 and the nine of their backward
// (autograd of A2C/algo/ppo.py:138-142).  Phases B (gradient reduce) and C (clip + Adam) are the ones of sg_ppo.cu.
//
// Job = (tile of MR minibatch rows, net) with net 0 = actor trunk + Gaussian mean head, 1 = critic trunk + critic_linear;
// CTA c runs jobs c, c + grid, ... (grid even, so a CTA's net is fixed).  One job is a chain of contractions
//   G1 Z1 = X W1^T   G2 Z2 = H1 W2^T   G3 head = H2 Wh^T   [per-row losses, seeds dHead]
//   G4 dWh^T = H2^T dHead   G5 dH2 = dHead Wh   G6 dW2 = dZ2^T H1   G7 dH1 = dZ2 W2   G8 dW1 = dZ1^T X
// each formed as D(tmem) = A.B^T by ONE thread issuing tcgen05.mma over K-chunks of shared-memory operand images.
//
// fp32 accuracy on TF32 tensor cores: every operand is split a = hi + lo (hi = RN_tf32(a), lo = RN_tf32(a - hi)) and a
// product is three MMAs lo*hi + hi*lo + hi*hi ("3xTF32", ~2^-21 relative), which keeps the 1e-4 loss contract.
//
// Data flow: activations live once, in fp32, in shared-memory "masters" laid out [col/4][row][col%4] (16-byte granules,
// row pitch MR+4 granules); weights stay in global memory (L2).  For every K-chunk all 256 threads build the hi / lo operand
// images of the chunk in one of two stage buffers (K-major or MN-major, sg_mma.cuh) while the tensor core works on the
// other one (tcgen05.commit -> mbarrier per stage); the next chunk's source values are already in registers when the
// current one is stored, so the L2 latency of the weight reads hides behind the MMAs.  Epilogues read the accumulator
// with tcgen05.ld (thread = row, or = hidden unit for the weight gradients), apply bias / tanh / tanh' and write the
// next master or the CTA's partial gradient.
#pragma once
#include "sg_common.cuh"
#include "sg_policy.cuh"
#include "sg_mma.cuh"

namespace sg {

constexpr int kMmaStageBytes = 24 * 1024;      // one stage buffer (hi + lo images of the A and B chunk); two of them
constexpr int kMmaRedStride = 36;

struct MmaDims {
    int MR, MRP;            // rows per job; master row pitch in granules
    int O, H, A;
    int Op8, Op32;          // K extent of layer 1; N extent of the dW1 contraction
    int xg, hg, dg;         // granule columns of the X / hidden / dHead masters
    int Mb, nblk;           // weight-gradient contractions: rows (hidden units) per MMA and number of such blocks
    int tmem_cols;
    // shared-memory carve-up (bytes from the 1024-aligned base)
    int o_stage, o_X, o_H1, o_H2, o_DH, o_small, total;
};

__host__ __device__ inline MmaDims make_mma_dims(int O, int H, int A, int MR) {
    MmaDims d;
    d.MR = MR; d.MRP = MR + 4;
    d.O = O; d.H = H; d.A = A;
    d.Op8 = round_up(O, 8); d.Op32 = round_up(O, 32);
    d.xg = d.Op8 / 4; d.hg = H / 4; d.dg = 8;
    d.Mb = H >= 128 ? 128 : 64;
    d.nblk = H / d.Mb;
    int cols = H > d.Op32 ? H : d.Op32;
    int t = 32;
    while (t < cols) t <<= 1;
    d.tmem_cols = t;
    int o = 0;
    d.o_stage = o; o += 2 * kMmaStageBytes;
    d.o_X = o; o += d.xg * d.MRP * 16;
    d.o_H1 = o; o += d.hg * d.MRP * 16;
    d.o_H2 = o; o += d.hg * d.MRP * 16;
    d.o_DH = o; o += d.dg * d.MRP * 16;
    d.o_small = o;
    // B1 B2 (H each), BH LS (32 each), 5 row-scalar arrays + IDX (MR each), GB1 GB2 (H), GBH GLS (32), LOSS (4),
    // RED (8 warps x 36), 2 mbarriers + tmem slot (32 bytes)
    o += (4 * H + 4 * 32 + 6 * MR + 4 + 8 * kMmaRedStride) * 4 + 32;
    d.total = round_up(o, 16);
    return d;
}

// H in {64,128,256}, A <= 32, O <= 256
__host__ inline bool ppo_mma_supported(int O, int H, int A) { return (H == 64 || H == 128 || H == 256) && A >= 1 && A <= 32 && O >= 1 && O <= 256; }

struct MmaSmem {
    float* stage[2];
    float4 *X, *H1, *H2, *DH;
    float *B1, *B2, *BH, *LS, *RET, *VP, *OLP, *ADV, *VALID, *GB1, *GB2, *GBH, *GLS, *LOSS, *RED;
    int* IDX;
    unsigned long long* bar;
    uint32_t* tmem_slot;
    __device__ void carve(unsigned char* base, const MmaDims& d) {
        stage[0] = reinterpret_cast<float*>(base + d.o_stage);
        stage[1] = reinterpret_cast<float*>(base + d.o_stage + kMmaStageBytes);
        X = reinterpret_cast<float4*>(base + d.o_X);
        H1 = reinterpret_cast<float4*>(base + d.o_H1);
        H2 = reinterpret_cast<float4*>(base + d.o_H2);
        DH = reinterpret_cast<float4*>(base + d.o_DH);
        float* f = reinterpret_cast<float*>(base + d.o_small);
        B1 = f; f += d.H;
        B2 = f; f += d.H;
        BH = f; f += 32;
        LS = f; f += 32;
        RET = f; f += d.MR;
        VP = f; f += d.MR;
        OLP = f; f += d.MR;
        ADV = f; f += d.MR;
        VALID = f; f += d.MR;
        IDX = reinterpret_cast<int*>(f); f += d.MR;
        GB1 = f; f += d.H;
        GB2 = f; f += d.H;
        GBH = f; f += 32;
        GLS = f; f += 32;
        LOSS = f; f += 4;
        RED = f; f += 8 * kMmaRedStride;
        bar = reinterpret_cast<unsigned long long*>(f);
        tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    }
};

// ---- operand sources -------------------------------------------------------------------------------------------------------
struct MmaOperand {
    int kind;                 // 0 K-major <- master, 1 MN-major <- master, 2 K-major <- global, 3 MN-major <- global
    int E;                    // MN extent of the stage image (multiple of 8; of 32 when MN-major)
    const float4* m4;         // master (kinds 0, 1)
    int e0, ecols;            // kind 1: first master column of the image and number of valid master columns
    const float* g;           // global row-major matrix (kinds 2, 3), read through L2 (rewritten by Adam between steps)
    int ld, nvalid, kvalid;   // row pitch in floats; valid MN extent; valid K extent
};
__device__ __forceinline__ MmaOperand op_master_k(const float4* m4, int E) { return MmaOperand{0, E, m4, 0, 0, nullptr, 0, 0, 0}; }
__device__ __forceinline__ MmaOperand op_master_mn(const float4* m4, int E, int e0, int ecols) {
    return MmaOperand{1, E, m4, e0, ecols, nullptr, 0, 0, 0};
}
// W is (nvalid, kvalid) row-major with pitch ld: element (e, k) = W[e*ld + k]
__device__ __forceinline__ MmaOperand op_global_k(const float* W, int E, int ld, int nvalid, int kvalid) {
    return MmaOperand{2, E, nullptr, 0, 0, W, ld, nvalid, kvalid};
}
// W is (kvalid, nvalid) row-major with pitch ld: element (e, k) = W[k*ld + e]
__device__ __forceinline__ MmaOperand op_global_mn(const float* W, int E, int ld, int nvalid, int kvalid) {
    return MmaOperand{3, E, nullptr, 0, 0, W, ld, nvalid, kvalid};
}

// granule gi of the chunk [k0, k0+kc): its four source floats and where they go in the stage image (float offset)
__device__ __forceinline__ float4 mma_src(const MmaOperand& op, int gi, int k0, int MRP, int& dst) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (op.kind == 0) {
        const int kq = gi / op.E, e = gi - kq * op.E;
        dst = gi << 2;
        return op.m4[((k0 >> 2) + kq) * MRP + e];
    }
    if (op.kind == 2) {
        const int kq = gi / op.E, e = gi - kq * op.E;
        dst = gi << 2;
        const int k = k0 + 4 * kq;
        if (e >= op.nvalid || k >= op.kvalid) return zero;
        const float* p = op.g + (size_t)e * op.ld + k;
        if ((op.ld & 3) == 0 && k + 3 < op.kvalid) return ld_cg4(p);
        float4 v = zero;
        v.x = ld_cg(p);
        if (k + 1 < op.kvalid) v.y = ld_cg(p + 1);
        if (k + 2 < op.kvalid) v.z = ld_cg(p + 2);
        if (k + 3 < op.kvalid) v.w = ld_cg(p + 3);
        return v;
    }
    // MN-major: lanes walk (kk = 4 k-rows, eq8 = 8 quads of one 32-element block): conflict-free stores of whole atoms
    const int kk = gi & 3, eq8 = (gi >> 2) & 7, t = gi >> 5;
    const int EB = op.E >> 5;
    const int kb = t / EB, eb = t - kb * EB;
    const int k = 4 * kb + kk, e = 32 * eb + 4 * eq8;
    dst = mma::mnmajor_off(e, k, op.E);
    if (op.kind == 1) {
        const int c = op.e0 + e;
        return c < op.ecols ? op.m4[(c >> 2) * MRP + k0 + k] : zero;
    }
    const int kg = k0 + k;
    if (kg >= op.kvalid || e >= op.nvalid) return zero;
    const float* p = op.g + (size_t)kg * op.ld + e;
    if ((op.ld & 3) == 0 && e + 3 < op.nvalid) return ld_cg4(p);
    float4 v = zero;
    v.x = ld_cg(p);
    if (e + 1 < op.nvalid) v.y = ld_cg(p + 1);
    if (e + 2 < op.nvalid) v.z = ld_cg(p + 2);
    if (e + 3 < op.nvalid) v.w = ld_cg(p + 3);
    return v;
}

constexpr int kMmaMaxG = 3;       // granules per thread, operand and chunk: 24 KiB / 8 bytes per element / 4 / 256 threads

struct MmaRegs {
    float4 v[kMmaMaxG];
    int dst[kMmaMaxG];
};
__device__ __forceinline__ void mma_load(const MmaOperand& op, int k0, int kc, int MRP, MmaRegs& r) {
    const int ng = (kc * op.E) >> 2;
#pragma unroll
    for (int i = 0; i < kMmaMaxG; ++i) {
        const int gi = threadIdx.x + i * kStepThreads;
        r.dst[i] = -1;
        if (gi < ng) r.v[i] = mma_src(op, gi, k0, MRP, r.dst[i]);
    }
}
__device__ __forceinline__ void mma_store(const MmaRegs& r, float* hi, float* lo) {
#pragma unroll
    for (int i = 0; i < kMmaMaxG; ++i) {
        if (r.dst[i] >= 0) {
            float4 h, l;
            mma::split4(r.v[i], h, l);
            *reinterpret_cast<float4*>(hi + r.dst[i]) = h;
            *reinterpret_cast<float4*>(lo + r.dst[i]) = l;
        }
    }
}

// ---- stage pipeline ------------------------------------------------------------------------------------------------------
struct MmaPipe {
    unsigned int par[2];
    bool pend[2];
    int cur;
    __device__ __forceinline__ void init() { par[0] = par[1] = 0u; pend[0] = pend[1] = false; cur = 0; }
    // the MMAs that read stage s have completed (its buffer may be rewritten)
    __device__ __forceinline__ void wait(unsigned long long* bar, int s) {
        if (pend[s]) { mbar_wait(bar + s, par[s]); par[s] ^= 1u; pend[s] = false; }
    }
};

// D[tmem columns dcol .. dcol+N) (Mi rows) = A . B^T over K, 3xTF32.  All 256 threads call; ends with every MMA complete
// and visible to tcgen05.ld (callers run their epilogue right away).
__device__ void mma_gemm(MmaSmem& S, MmaPipe& P, int MRP, uint32_t tbase, uint32_t dcol, int Mi, int N, int K,
                         const MmaOperand& A, const MmaOperand& B) {
    const int tid = threadIdx.x;
    const int a_mn = A.kind & 1, b_mn = B.kind & 1;
    int KC = (kMmaStageBytes / ((A.E + B.E) * 8)) & ~7;
    if (KC > K) KC = K;
    const uint32_t idesc = mma::make_idesc_tf32(Mi, N, a_mn, b_mn);
    const uint32_t a_lbo = a_mn ? 512u : 16u * A.E, a_sbo = a_mn ? 16u * A.E : 128u;
    const uint32_t b_lbo = b_mn ? 512u : 16u * B.E, b_sbo = b_mn ? 16u * B.E : 128u;
    MmaRegs ra, rb;
    mma_load(A, 0, KC < K ? KC : K, MRP, ra);
    mma_load(B, 0, KC < K ? KC : K, MRP, rb);
    uint32_t accum = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = K - k0 < KC ? K - k0 : KC;
        const int s = P.cur;
        P.cur ^= 1;
        P.wait(S.bar, s);
        float* Ahi = S.stage[s];
        float* Alo = Ahi + A.E * kc;
        float* Bhi = Alo + A.E * kc;
        float* Blo = Bhi + B.E * kc;
        mma_store(ra, Ahi, Alo);
        mma_store(rb, Bhi, Blo);
        if (k0 + kc < K) {
            const int kn = K - (k0 + kc) < KC ? K - (k0 + kc) : KC;
            mma_load(A, k0 + kc, kn, MRP, ra);
            mma_load(B, k0 + kc, kn, MRP, rb);
        }
        mma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            mma::fence_after_sync();
            const uint32_t ah = mma::smem_addr(Ahi), al = mma::smem_addr(Alo), bh = mma::smem_addr(Bhi), bl = mma::smem_addr(Blo);
            for (int ks = 0; ks < (kc >> 3); ++ks) {
                const uint32_t sa = (uint32_t)ks * 32u * A.E, sb = (uint32_t)ks * 32u * B.E;
                const uint64_t dAh = mma::make_desc(ah + sa, a_lbo, a_sbo, a_mn), dAl = mma::make_desc(al + sa, a_lbo, a_sbo, a_mn);
                const uint64_t dBh = mma::make_desc(bh + sb, b_lbo, b_sbo, b_mn), dBl = mma::make_desc(bl + sb, b_lbo, b_sbo, b_mn);
                mma::mma_tf32(tbase + dcol, dAl, dBh, idesc, accum);
                mma::mma_tf32(tbase + dcol, dAh, dBl, idesc, 1u);
                mma::mma_tf32(tbase + dcol, dAh, dBh, idesc, 1u);
                accum = 1u;
            }
            mma::commit(S.bar + s);
        }
        P.pend[s] = true;
    }
    P.wait(S.bar, 0);
    P.wait(S.bar, 1);
    mma::fence_after_sync();
}

// Epilogue walker: accumulator rows [0, Mi) x columns [0, N) from tmem column dcol; f(m, c0, v) gets 16 consecutive
// columns c0.. of row m.  Warps w and w+4 share a lane quarter and alternate 16-column chunks.  Ends with a CTA barrier.
template <class F>
__device__ __forceinline__ void mma_epilogue(uint32_t tbase, uint32_t dcol, int Mi, int N, F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int m = Mi == 128 ? 32 * q + lane : 16 * q + lane;
    const bool live = Mi == 128 || lane < 16;
    for (int c0 = 16 * half; c0 < N; c0 += 32) {
        float v[16];
        mma::tmem_ld16(tbase + ((uint32_t)(32 * q) << 16) + dcol + (uint32_t)c0, v);
        mma::tmem_ld_wait();
        if (live) f(m, c0, v);
    }
    mma::fence_before_sync();
    __syncthreads();
}

// acc[c] += sum over rows of master column c (bias gradients); deterministic; ends with a CTA barrier
__device__ __forceinline__ void mma_colsum(const float4* m4, int MRP, int MR, int ng, float* acc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cq = warp; cq < ng; cq += kStepThreads / 32) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = lane; r < MR; r += 32) s = f4_add(s, m4[cq * MRP + r]);
        s.x = warp_sum(s.x); s.y = warp_sum(s.y); s.z = warp_sum(s.z); s.w = warp_sum(s.w);
        if (lane == 0) {
            float* p = acc + 4 * cq;
            p[0] += s.x; p[1] += s.y; p[2] += s.z; p[3] += s.w;
        }
    }
    __syncthreads();
}

// ---- one job ------------------------------------------------------------------------------------------------------------------
template <int MR>
__device__ void ppo_mma_job(const PpoArgs& a, const MmaDims& d, MmaSmem& S, MmaPipe& P, uint32_t tbase, int step, int tile,
                            int net, float* __restrict__ gout, bool acc) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int O = a.O, H = a.H, A = a.A, MRP = d.MRP;
    const int epoch = step / a.nmb, mb = step - epoch * a.nmb;
    const int32_t* idx = a.perm + (size_t)epoch * a.S + (size_t)mb * a.mbs;
    const int row0 = a.row_begin + tile * MR;
    const PolicyLayout& L = a.L;
    const float* W1 = a.params + (net ? L.cw1 : L.aw1);
    const float* W2 = a.params + (net ? L.cw2 : L.aw2);
    const float* Wh = a.params + (net ? L.vw : L.mw);
    const int NA = net ? 1 : A;
    const int NA16 = round_up(NA, 16), NA8 = round_up(NA, 8);

    // ---- sampler indices and row scalars (flat sample id = t*N+n, A2C/storage.py:169-185) -------------------------------
    if (tid < MR) {
        const int row = row0 + tid;
        const bool ok = row < a.row_end;
        const int i = ok ? idx[row] : -1;
        S.IDX[tid] = i;
        const float ret = ok ? a.ret[i] : 0.f, vp = ok ? a.vpred[i] : 0.f;
        S.RET[tid] = ret; S.VP[tid] = vp; S.OLP[tid] = ok ? a.oldlp[i] : 0.f;
        const float mean = a.advstats[0], sd = a.advstats[1];
        S.ADV[tid] = ok ? __fdiv_rn(__fsub_rn(__fsub_rn(ret, vp), mean), __fadd_rn(sd, 1e-5f)) : 0.f;   // ppo.py:66-68
        S.VALID[tid] = ok ? 1.f : 0.f;
    }
    __syncthreads();
    // ---- gather the observation rows into the X master: a warp covers 8 rows x 4 granules (64 contiguous bytes per row)
    {
        const int kq_l = lane & 3, r_l = lane >> 2;
        const int kgroups = (d.xg + 3) >> 2;
        const bool vec = (O & 3) == 0;
        for (int it = warp; it < (MR / 8) * kgroups; it += kStepThreads / 32) {
            const int rg = it / kgroups, kg = it - rg * kgroups;
            const int r = 8 * rg + r_l, kq = 4 * kg + kq_l;
            if (kq < d.xg) {
                const int i = S.IDX[r];
                const int k = 4 * kq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i >= 0 && k < O) {
                    const float* p = a.obs + (size_t)i * O + k;
                    if (vec) v = *reinterpret_cast<const float4*>(p);
                    else {
                        v.x = p[0];
                        if (k + 1 < O) v.y = p[1];
                        if (k + 2 < O) v.z = p[2];
                        if (k + 3 < O) v.w = p[3];
                    }
                }
                S.X[kq * MRP + r] = v;
            }
        }
    }
    __syncthreads();

    // ---- forward (A2C/model.py:255-264) ---------------------------------------------------------------------------------------
    mma_gemm(S, P, MRP, tbase, 0, MR, H, d.Op8, op_master_k(S.X, MR), op_global_k(W1, H, O, H, O));
    mma_epilogue(tbase, 0, MR, H, [&](int r, int c0, const float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            S.H1[(c >> 2) * MRP + r] = make_float4(tanhf(v[4 * j] + S.B1[c]), tanhf(v[4 * j + 1] + S.B1[c + 1]),
                                                   tanhf(v[4 * j + 2] + S.B1[c + 2]), tanhf(v[4 * j + 3] + S.B1[c + 3]));
        }
    });
    mma_gemm(S, P, MRP, tbase, 0, MR, H, H, op_master_k(S.H1, MR), op_global_k(W2, H, H, H, H));
    mma_epilogue(tbase, 0, MR, H, [&](int r, int c0, const float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            S.H2[(c >> 2) * MRP + r] = make_float4(tanhf(v[4 * j] + S.B2[c]), tanhf(v[4 * j + 1] + S.B2[c + 1]),
                                                   tanhf(v[4 * j + 2] + S.B2[c + 2]), tanhf(v[4 * j + 3] + S.B2[c + 3]));
        }
    });
    // head: Gaussian mean (A2C/distributions.py:109-110) or critic_linear
    mma_gemm(S, P, MRP, tbase, 0, MR, NA16, H, op_master_k(S.H2, MR), op_global_k(Wh, NA16, H, NA, H));

    // ---- per-row losses and the gradient seeds (thread = row; warps 0-3) -------------------------------------------------------
    {
        const int q = warp & 3;
        const int r = MR == 128 ? 32 * q + lane : 16 * q + lane;
        const bool live = warp < 4 && (MR == 128 || lane < 16);
        float vl = 0.f, al = 0.f;
        if (warp < 4) {
            float v[32];
            {
                float t16[16];
                mma::tmem_ld16(tbase + ((uint32_t)(32 * q) << 16), t16);
                mma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = t16[j];
                if (NA16 > 16) {
                    mma::tmem_ld16(tbase + ((uint32_t)(32 * q) << 16) + 16u, t16);
                    mma::tmem_ld_wait();
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[16 + j] = NA16 > 16 ? t16[j] : 0.f;
            }
            const bool ok = live && S.VALID[r] != 0.f;
            const float invB = 1.f / (float)a.mbs;
            if (net == 0) {
                const int i = live ? S.IDX[r] : -1;
                const float* act = a.actions + (size_t)(i >= 0 ? i : 0) * A;
                // log-prob of the stored action, summed over the action dim (A2C/distributions.py:52-53); v[k] <- a - mu
                float lp = 0.f;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    if (k < A) {
                        const float sigma = expf(S.LS[k]);
                        const float var = sigma * sigma;
                        const float dd = (ok ? act[k] : 0.f) - (v[k] + S.BH[k]);
                        v[k] = dd;
                        lp += -(dd * dd) / (2.f * var) - logf(sigma) - SG_LOG_SQRT_2PI;
                    }
                }
                float coef = 0.f;
                if (ok) {
                    const float ratio = expf(lp - S.OLP[r]);
                    const float adv = S.ADV[r];
                    const float s1 = ratio * adv;
                    const float s2 = fminf(fmaxf(ratio, a.ratio_lo), a.ratio_hi) * adv;
                    al = -fminf(s1, s2);
                    const float inr = (ratio >= a.ratio_lo && ratio <= a.ratio_hi) ? 1.f : 0.f;
                    // torch.min backward: all to the smaller side, 1/2 + 1/2 on exact ties; clamp passes grad on inclusive bounds
                    const float gsel = s1 < s2 ? 1.f : (s2 < s1 ? inr : 0.5f + 0.5f * inr);
                    coef = -invB * gsel * adv * ratio;
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    float dmu = 0.f, dls = 0.f;
                    if (k < A) {
                        const float sigma = expf(S.LS[k]);
                        const float var = sigma * sigma;
                        const float dd = v[k];
                        dmu = ok ? coef * dd / var : 0.f;                 // d logp / d mu     = (a-mu)/var
                        dls = ok ? coef * (dd * dd / var - 1.f) : 0.f;    // d logp / d logstd = (a-mu)^2/var - 1
                        dls = warp_sum(dls);
                        if (lane == 0) S.RED[warp * kMmaRedStride + k] = dls;
                    }
                    v[k] = dmu;
                }
            } else {
                float dv = 0.f;
                if (ok) {
                    const float val = v[0] + S.BH[0], vp = S.VP[r], ret = S.RET[r];
                    if (a.clipped_vloss) {
                        const float diff = val - vp;
                        const float vc = vp + fminf(fmaxf(diff, -a.clip), a.clip);
                        const float e1 = val - ret, e2 = vc - ret;
                        const float l1 = e1 * e1, l2 = e2 * e2;
                        vl = 0.5f * fmaxf(l1, l2);
                        const float in2 = (diff >= -a.clip && diff <= a.clip) ? 1.f : 0.f;
                        const float g = l1 > l2 ? e1 : (l2 > l1 ? in2 * e2 : 0.5f * (e1 + in2 * e2));
                        dv = a.c_v * invB * g;
                    } else {
                        const float e1 = ret - val;
                        vl = 0.5f * e1 * e1;
                        dv = a.c_v * invB * (val - ret);
                    }
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = 0.f;
                v[0] = dv;
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < 8; ++j) S.DH[j * MRP + r] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            vl = warp_sum(vl);
            al = warp_sum(al);
            if (lane == 0) { S.RED[warp * kMmaRedStride + 32] = vl; S.RED[warp * kMmaRedStride + 33] = al; }
        }
        mma::fence_before_sync();
        __syncthreads();
        if (tid < 34) {
            const float s = S.RED[tid] + S.RED[kMmaRedStride + tid] + S.RED[2 * kMmaRedStride + tid] + S.RED[3 * kMmaRedStride + tid];
            if (tid < 32) { if (net == 0 && tid < A) S.GLS[tid] += s; }
            else S.LOSS[tid - 32] += s;
        }
        mma_colsum(S.DH, MRP, MR, d.dg, S.GBH);
    }

    // ---- backward ------------------------------------------------------------------------------------------------------------------
    // head weight gradient, transposed: D(unit, a) = sum_rows H2(row, unit) dHead(row, a)
    for (int b = 0; b < d.nblk; ++b) {
        mma_gemm(S, P, MRP, tbase, 0, d.Mb, 32, MR, op_master_mn(S.H2, d.Mb, b * d.Mb, H), op_master_mn(S.DH, 32, 0, 32));
        float* gW = gout + (net ? L.vw : L.mw);
        mma_epilogue(tbase, 0, d.Mb, 32, [&](int m, int c0, const float (&v)[16]) {
            const int u = b * d.Mb + m;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = c0 + j;
                if (k < NA) {
                    float* p = gW + (size_t)k * H + u;
                    __stcg(p, acc ? v[j] + __ldcg(p) : v[j]);
                }
            }
        });
    }
    // dZ2 = (dHead . Wh) * (1 - h2^2), in place over the H2 master
    mma_gemm(S, P, MRP, tbase, 0, MR, H, NA8, op_master_k(S.DH, MR), op_global_mn(Wh, H, H, H, NA));
    mma_epilogue(tbase, 0, MR, H, [&](int r, int c0, const float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4* p = S.H2 + ((c0 >> 2) + j) * MRP + r;
            const float4 h = *p;
            *p = make_float4(v[4 * j] * (1.f - h.x * h.x), v[4 * j + 1] * (1.f - h.y * h.y), v[4 * j + 2] * (1.f - h.z * h.z),
                             v[4 * j + 3] * (1.f - h.w * h.w));
        }
    });
    mma_colsum(S.H2, MRP, MR, d.hg, S.GB2);
    // dW2(n, k) = sum_rows dZ2(row, n) H1(row, k)
    for (int b = 0; b < d.nblk; ++b) {
        mma_gemm(S, P, MRP, tbase, 0, d.Mb, H, MR, op_master_mn(S.H2, d.Mb, b * d.Mb, H), op_master_mn(S.H1, H, 0, H));
        float* gW = gout + (net ? L.cw2 : L.aw2);
        mma_epilogue(tbase, 0, d.Mb, H, [&](int m, int c0, const float (&v)[16]) {
            float* p = gW + (size_t)(b * d.Mb + m) * H + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                float4* p4 = reinterpret_cast<float4*>(p) + j;
                if (acc) o = f4_add(o, __ldcg(p4));
                __stcg(p4, o);
            }
        });
    }
    // dZ1 = (dZ2 . W2) * (1 - h1^2), in place over the H1 master
    mma_gemm(S, P, MRP, tbase, 0, MR, H, H, op_master_k(S.H2, MR), op_global_mn(W2, H, H, H, H));
    mma_epilogue(tbase, 0, MR, H, [&](int r, int c0, const float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4* p = S.H1 + ((c0 >> 2) + j) * MRP + r;
            const float4 h = *p;
            *p = make_float4(v[4 * j] * (1.f - h.x * h.x), v[4 * j + 1] * (1.f - h.y * h.y), v[4 * j + 2] * (1.f - h.z * h.z),
                             v[4 * j + 3] * (1.f - h.w * h.w));
        }
    });
    mma_colsum(S.H1, MRP, MR, d.hg, S.GB1);
    // dW1(n, k) = sum_rows dZ1(row, n) X(row, k)
    for (int b = 0; b < d.nblk; ++b) {
        mma_gemm(S, P, MRP, tbase, 0, d.Mb, d.Op32, MR, op_master_mn(S.H1, d.Mb, b * d.Mb, H), op_master_mn(S.X, d.Op32, 0, d.Op8));
        float* gW = gout + (net ? L.cw1 : L.aw1);
        const bool vec = (O & 3) == 0;
        mma_epilogue(tbase, 0, d.Mb, d.Op32, [&](int m, int c0, const float (&v)[16]) {
            float* p = gW + (size_t)(b * d.Mb + m) * O + c0;
            if (vec) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c0 + 4 * j < O) {
                        float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        float4* p4 = reinterpret_cast<float4*>(p) + j;
                        if (acc) o = f4_add(o, __ldcg(p4));
                        __stcg(p4, o);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < O) __stcg(p + j, acc ? v[j] + __ldcg(p + j) : v[j]);
            }
        });
    }
}

// ---- phase A of one optimizer step -------------------------------------------------------------------------------------------
template <int MR>
__device__ void ppo_phaseA_mma(const PpoArgs& a, const MmaDims& d, MmaSmem& S, MmaPipe& P, uint32_t tbase, int step, int cta, int ncta) {
    const int tid = threadIdx.x;
    const int njobs = 2 * a.ntiles;
    if (cta >= njobs) return;
    const int net = cta & 1;
    const PolicyLayout& L = a.L;
    const int H = a.H, A = a.A, NA = net ? 1 : A;
    // this step's biases / log-std of my net; zero the per-step accumulators
    for (int i = tid; i < H; i += kStepThreads) {
        S.B1[i] = ld_cg(a.params + (net ? L.cb1 : L.ab1) + i);
        S.B2[i] = ld_cg(a.params + (net ? L.cb2 : L.ab2) + i);
        S.GB1[i] = 0.f; S.GB2[i] = 0.f;
    }
    if (tid < 32) {
        S.BH[tid] = tid < NA ? ld_cg(a.params + (net ? L.vb : L.mb) + tid) : 0.f;
        S.LS[tid] = tid < A ? ld_cg(a.params + L.ls + tid) : 0.f;
        S.GBH[tid] = 0.f; S.GLS[tid] = 0.f;
        if (tid < 4) S.LOSS[tid] = 0.f;
    }
    __syncthreads();
    float* gout = a.gpart + (size_t)cta * a.P;
    bool acc = false;
    for (int job = cta; job < njobs; job += ncta) {
        ppo_mma_job<MR>(a, d, S, P, tbase, step, job >> 1, net, gout, acc);
        acc = true;
    }
    for (int i = tid; i < H; i += kStepThreads) {
        __stcg(gout + (net ? L.cb1 : L.ab1) + i, S.GB1[i]);
        __stcg(gout + (net ? L.cb2 : L.ab2) + i, S.GB2[i]);
    }
    if (tid < NA) __stcg(gout + (net ? L.vb : L.mb) + tid, S.GBH[tid]);
    if (net == 0 && tid < A) __stcg(gout + L.ls + tid, S.GLS[tid]);
    if (tid == 0) { a.losspart[cta * 4] = S.LOSS[0]; a.losspart[cta * 4 + 1] = S.LOSS[1]; }
}

}  // namespace sg
