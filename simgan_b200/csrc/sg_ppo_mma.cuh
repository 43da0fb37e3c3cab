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
// fp32 accuracy on TF32 tensor cores: every operand is split a = hi + lo (activations: hi = the top 19 bits of the word, which
// is what the tensor core reads of it anyway, lo = a - hi exactly; weight images: hi = RN_tf32(a), lo = RN_tf32(a - hi)) and a
// product is three MMAs lo*hi + hi*lo + hi*hi ("3xTF32", ~2^-21 relative), which keeps the 1e-4 loss contract.
//
// Data flow
//   activations  live once, in fp32, in shared-memory "masters" laid out [col/4][row][col%4] (16-byte granules, row
//                pitch MR+1 granules).  That IS the K-major operand layout, and the tensor core reads only the top 19 bits
//                of an fp32 word, so a K-major activation operand's hi pass reads the master in place; per K-chunk the
//                seven "filler" warps build only its lo image (x - trunc(x)) -- or, for the MN-major operands of the
//                weight-gradient contractions, the hi and lo images -- in a ring of stage buffers (sg_mma.cuh);
//   pipeline     warp 0 is the control warp: one thread issues the MMAs; the first filler thread requests the weight chunks (a
//                stage is refilled by whoever sees it free, so the issuer never waits for a completion); fillers and issuer
//                meet only through mbarriers (filled[s]: image ready, full[s]: weight chunk landed, done[s]: the MMAs that
//                read stage s completed), so filling chunk c+1.., the TMA of chunk c+2.. and the MMAs of chunk c overlap;
//                the CTA only joins again at the epilogue;
//   weights      are kept as ready-made hi / lo operand images of the whole K extent in global memory (L2): the CTA that
//                owns a slice of the parameter vector rewrites the image entries of its parameters right after their Adam
//                update, so a K-chunk of a weight operand is two contiguous TMA bulk copies (cp.async.bulk -> mbarrier
//                transaction bytes) issued as many chunks ahead as the ring has stages -- no register staging, no split
//                arithmetic;
//   accumulators live in tensor memory; epilogues read them with tcgen05.ld (thread = row, or = hidden unit for the weight
//                gradients), apply bias / tanh / tanh' and write the next master or the CTA's partial gradient.  When the
//                three weight-gradient accumulators fit beside the working accumulator (hidden <= 128) they stay in
//                tensor memory across all jobs of the CTA and are flushed once per optimizer step.
#pragma once
#include "sg_common.cuh"
#include "sg_policy.cuh"
#include "sg_mma.cuh"
#ifdef SG_MMA_PROFILE
#include <stdio.h>
#endif

namespace sg {

constexpr int kMmaRedStride = 36;
constexpr int kMmaMaxStages = 6;
constexpr int kMmaFillThreads = kStepThreads - 32;      // warps 1.. build operand images; warp 0 issues TMA + MMA
constexpr int kMmaMinRingBytes = 40 * 1024, kMmaMaxRingBytes = 96 * 1024;

struct MmaDims {
    int MR, MRP;            // rows per job; master row pitch in granules
    int O, H, A;
    int Op8, Op32;          // K extent of layer 1; N extent of the dW1 contraction
    int xg, hg, dg, ag;     // granule columns of the X / hidden / dHead / action masters
    int Mb, nblk;           // weight-gradient contractions: rows (hidden units) per MMA and number of such blocks
    int ring;               // bytes of the stage ring (carved per contraction into NS stages of one K-chunk each)
    int tacc;               // weight-gradient accumulators stay in tensor memory across the CTA's jobs
    int c_dwh, c_dw2, c_dw1;// their columns (tacc)
    int tmem_cols;
    // shared-memory carve-up (bytes from the 1024-aligned base)
    int o_stage, o_X, o_ACT, o_H1, o_H2, o_DH, o_small, total;
};

__host__ __device__ inline MmaDims make_mma_dims(int O, int H, int A, int MR, int ring_bytes) {
    MmaDims d;
    d.MR = MR; d.MRP = MR + 1;
    d.O = O; d.H = H; d.A = A;
    d.Op8 = round_up(O, 8); d.Op32 = round_up(O, 32);
    d.xg = d.Op8 / 4; d.hg = H / 4; d.dg = 8; d.ag = round_up(A, 4) / 4;
    d.Mb = H >= 128 ? 128 : 64;
    d.nblk = H / d.Mb;
    d.ring = ring_bytes;
    const int work = H > 32 ? H : 32;
    d.c_dwh = work; d.c_dw2 = work + 32; d.c_dw1 = work + 32 + H;
    d.tacc = (d.nblk == 1 && d.c_dw1 + d.Op32 <= 512) ? 1 : 0;
    int cols = d.tacc ? d.c_dw1 + d.Op32 : (H > d.Op32 ? H : d.Op32);
    int t = 32;
    while (t < cols) t <<= 1;
    d.tmem_cols = t;
    int o = 0;
    d.o_stage = o; o += ring_bytes;
    d.o_X = o; o += d.xg * d.MRP * 16;
    d.o_ACT = o; o += d.ag * d.MRP * 16;
    d.o_H1 = o; o += d.hg * d.MRP * 16;
    d.o_H2 = o; o += d.hg * d.MRP * 16;
    d.o_DH = o; o += d.dg * d.MRP * 16;
    d.o_small = o;
    // B1 B2 GB1 GB2 (H each), BH LS SIG VAR LSG GBH GLS (32 each), 5 row-scalar arrays + IDX (MR each), LOSS (4),
    // RED (8 warps x 36), 3 x kMmaMaxStages mbarriers + tmem slot (192 bytes)
    o += (4 * H + 7 * 32 + 6 * MR + 4 + 8 * kMmaRedStride) * 4 + 192;
    d.total = round_up(o, 16);
    return d;
}

// H in {64,128,256}, A <= 32, O <= 256
__host__ inline bool ppo_mma_supported(int O, int H, int A) { return (H == 64 || H == 128 || H == 256) && A >= 1 && A <= 32 && O >= 1 && O <= 256; }

// ---- weight operand images in global memory -----------------------------------------------------------------------------
// per net, hi image then lo image (net_floats apart): W1 K-major (E = H, K = Op8), W2 K-major (E = H, K = H), head K-major
// (E = NA16, K = H), head MN-major (E = H, K = NA8), W2 MN-major (E = H, K = H); zero where padded.
struct MmaImg {
    int w1k, w2k, whk, whm, w2m, net_floats;
};
__host__ __device__ inline MmaImg make_mma_img(int O, int H) {
    MmaImg g;
    int o = 0;
    g.w1k = o; o += H * round_up(O, 8);
    g.w2k = o; o += H * H;
    g.whk = o; o += 32 * H;
    g.whm = o; o += 32 * H;
    g.w2m = o; o += H * H;
    g.net_floats = o;
    return g;
}
__host__ __device__ inline size_t mma_img_total_floats(int O, int H) { return (size_t)4 * make_mma_img(O, H).net_floats; }

// image entries of parameter p (flat index) <- w.  Called by the owner of the parameter slice after Adam.
__device__ __forceinline__ void mma_img_put(float* __restrict__ wimg, const MmaImg& g, const PolicyLayout& L, int O, int H, int A,
                                            int p, float w) {
    int net, off0, off1 = -1;
    if (p >= L.aw1 && p < L.aw1 + H * O) { net = 0; const int i = p - L.aw1, n = i / O, k = i - n * O; off0 = g.w1k + mma::kmajor_off(n, k, H); }
    else if (p >= L.cw1 && p < L.cw1 + H * O) { net = 1; const int i = p - L.cw1, n = i / O, k = i - n * O; off0 = g.w1k + mma::kmajor_off(n, k, H); }
    else if (p >= L.aw2 && p < L.aw2 + H * H) {
        net = 0; const int i = p - L.aw2, n = i / H, k = i - n * H;
        off0 = g.w2k + mma::kmajor_off(n, k, H); off1 = g.w2m + mma::mnmajor_off(k, n, H);
    } else if (p >= L.cw2 && p < L.cw2 + H * H) {
        net = 1; const int i = p - L.cw2, n = i / H, k = i - n * H;
        off0 = g.w2k + mma::kmajor_off(n, k, H); off1 = g.w2m + mma::mnmajor_off(k, n, H);
    } else if (p >= L.mw && p < L.mw + A * H) {
        net = 0; const int i = p - L.mw, n = i / H, k = i - n * H;
        off0 = g.whk + mma::kmajor_off(n, k, round_up(A, 16)); off1 = g.whm + mma::mnmajor_off(k, n, H);
    } else if (p >= L.vw && p < L.vw + H) {
        net = 1; const int k = p - L.vw;
        off0 = g.whk + mma::kmajor_off(0, k, 16); off1 = g.whm + mma::mnmajor_off(k, 0, H);
    } else return;
    float hi, lo;
    mma::split_tf32(w, hi, lo);
    float* base = wimg + (size_t)(2 * net) * g.net_floats;
    __stcg(base + off0, hi);
    __stcg(base + g.net_floats + off0, lo);
    if (off1 >= 0) { __stcg(base + off1, hi); __stcg(base + g.net_floats + off1, lo); }
}
// refresh the images of the parameter range [p0, p1) from the (just updated) flat vector; all threads of the CTA call
__device__ __forceinline__ void mma_img_refresh(float* __restrict__ wimg, const float* __restrict__ params, const PolicyLayout& L,
                                                int O, int H, int A, int p0, int p1) {
    const MmaImg g = make_mma_img(O, H);
    for (int p = p0 + threadIdx.x; p < p1; p += kStepThreads) mma_img_put(wimg, g, L, O, H, A, p, ld_cg(params + p));
}

struct MmaSmem {
    float* stage0;
    float4 *X, *ACT, *H1, *H2, *DH;
    float *B1, *B2, *GB1, *GB2, *BH, *LS, *IVAR, *I2VAR, *LSG, *GBH, *GLS, *RET, *VP, *OLP, *ADV, *VALID, *LOSS, *RED;
    int* IDX;
    unsigned long long *full, *done, *filled;
    uint32_t* tmem_slot;
    __device__ void carve(unsigned char* base, const MmaDims& d) {
        stage0 = reinterpret_cast<float*>(base + d.o_stage);
        X = reinterpret_cast<float4*>(base + d.o_X);
        ACT = reinterpret_cast<float4*>(base + d.o_ACT);
        H1 = reinterpret_cast<float4*>(base + d.o_H1);
        H2 = reinterpret_cast<float4*>(base + d.o_H2);
        DH = reinterpret_cast<float4*>(base + d.o_DH);
        float* f = reinterpret_cast<float*>(base + d.o_small);
        B1 = f; f += d.H;
        B2 = f; f += d.H;
        GB1 = f; f += d.H;
        GB2 = f; f += d.H;
        BH = f; f += 32;
        LS = f; f += 32;
        IVAR = f; f += 32;
        I2VAR = f; f += 32;
        LSG = f; f += 32;
        GBH = f; f += 32;
        GLS = f; f += 32;
        RET = f; f += d.MR;
        VP = f; f += d.MR;
        OLP = f; f += d.MR;
        ADV = f; f += d.MR;
        VALID = f; f += d.MR;
        IDX = reinterpret_cast<int*>(f); f += d.MR;
        LOSS = f; f += 4;
        RED = f; f += 8 * kMmaRedStride;
        full = reinterpret_cast<unsigned long long*>(f);
        done = full + kMmaMaxStages;
        filled = done + kMmaMaxStages;
        tmem_slot = reinterpret_cast<uint32_t*>(filled + kMmaMaxStages);
    }
};

// ---- operand sources -------------------------------------------------------------------------------------------------------
struct MmaOperand {
    int kind;                 // 0 K-major <- master, 1 MN-major <- master, 2 K-major <- global image (TMA), 3 MN-major image
    int E;                    // MN extent of the stage image (multiple of 8; of 32 when MN-major)
    const float4* m4;         // master (kinds 0, 1)
    int e0, ecols;            // kind 1: first master column of the image and number of valid master columns
    const float* img;         // kinds 2, 3: hi image of the whole K extent in global memory; lo image img_lo floats further
    int img_lo;
};
__device__ __forceinline__ MmaOperand op_master_k(const float4* m4, int E) { return MmaOperand{0, E, m4, 0, 0, nullptr, 0}; }
__device__ __forceinline__ MmaOperand op_master_mn(const float4* m4, int E, int e0, int ecols) { return MmaOperand{1, E, m4, e0, ecols, nullptr, 0}; }
__device__ __forceinline__ MmaOperand op_image(const float* img, int lo, int E, int mn) { return MmaOperand{2 + mn, E, nullptr, 0, 0, img, lo}; }

constexpr int kMmaMaxG = 4;       // granules per filler thread, operand and chunk (a filled operand chunk has <= 896 granules)

// per-thread, per-contraction plan of a master operand: which master granule each of my (up to 4) stage granules comes
// from at k0 = 0 and where it goes; both are independent of the chunk, so a chunk only adds an offset
struct MmaPlan {
    int src[kMmaMaxG], dst[kMmaMaxG];      // master granule index at k0 = 0 (-1: zero fill), stage float offset
};
__device__ __forceinline__ void mma_plan(const MmaOperand& op, int MRP, int ftid, MmaPlan& pl) {
#pragma unroll
    for (int i = 0; i < kMmaMaxG; ++i) {
        const int gi = ftid + i * kMmaFillThreads;
        if (op.kind == 0) {
            const int kq = gi / op.E, e = gi - kq * op.E;
            pl.src[i] = kq * MRP + e;
            pl.dst[i] = gi << 2;
        } else {
            // MN-major: a warp covers one atom (4 k-rows x 32 elements); the 8 lanes of a quarter-warp share the k-row and take
            // the 8 element quads: their master granules (c/4)*MRP + k are 8 different 16-byte bank groups (MRP is odd) and so
            // are their swizzled destinations -- conflict-free on both sides (lanes walking k fastest gave 2-way conflicts on
            // the loads)
            const int kk = (gi >> 3) & 3, eq8 = gi & 7, t = gi >> 5;
            const int EB = op.E >> 5;
            const int kb = t / EB, eb = t - kb * EB;
            const int k = 4 * kb + kk, e = 32 * eb + 4 * eq8;
            const int c = op.e0 + e;
            pl.src[i] = c < op.ecols ? (c >> 2) * MRP + k : -1;
            pl.dst[i] = mma::mnmajor_off(e, k, op.E);
        }
    }
}
// chunk [k0, k0+kc) of a master operand -> stage images.  hi = x with the 13 low mantissa bits cleared (what the tensor
// core reads of an fp32 word anyway), lo = x - hi exactly (13 significant bits, of which the tensor core keeps 11):
// x = hi + lo to 2^-21 relative, two instructions per element.  K-major operands (kind 0) get only the lo image: their hi
// pass reads the master itself.
__device__ __forceinline__ void mma_fill(const MmaOperand& op, const MmaPlan& pl, int MRP, int ftid, int k0, int kc, float* hi, float* lo) {
    const int ng = (kc * op.E) >> 2;
    const int delta = op.kind == 0 ? (k0 >> 2) * MRP : k0;
#pragma unroll
    for (int i = 0; i < kMmaMaxG; ++i) {
        if (i * kMmaFillThreads >= ng) break;
        const int gi = ftid + i * kMmaFillThreads;
        if (gi < ng) {
            const float4 x = pl.src[i] >= 0 ? op.m4[pl.src[i] + delta] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 h;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            if (op.kind != 0) *reinterpret_cast<float4*>(hi + pl.dst[i]) = h;
            *reinterpret_cast<float4*>(lo + pl.dst[i]) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
    }
}

// ---- stage pipeline ------------------------------------------------------------------------------------------------------
// Three mbarriers per stage: filled[s] (one arrival per filler warp: the chunk's images are written and fenced), full[s]
// (TMA transaction bytes of the chunk's weight images), done[s] (tcgen05.commit: the MMAs that read stage s completed, the
// stage may be rewritten).  A contraction starts with every stage free (its predecessor drained), so chunk c uses stage
// c % NS for the (c / NS)-th time and the parity of its phase on each barrier is base bit ^ (c / NS): every thread can
// compute any chunk's parity without having waited on the earlier ones.  The base bits advance at the end of a contraction,
// identically in all threads.
struct MmaPipe {
    unsigned int done_base, fill_base, full_base;      // bit s
#ifdef SG_MMA_PROFILE
    int gid;                                            // contraction index within the job (profile rows)
#endif
    __device__ __forceinline__ void init() {
        done_base = fill_base = full_base = 0u;
#ifdef SG_MMA_PROFILE
        gid = 0;
#endif
    }
};
#ifdef SG_MMA_PROFILE
// per contraction: {issuer: wait filled, wait full, issue, wait refill; filler (thread 32): wait done, fill; chunks, NS}
__shared__ unsigned int sg_mma_prof[20][10];
#define SG_MMA_CLK(v) const long long v = clock64()
#define SG_MMA_ADD(col, expr) sg_mma_prof[P.gid][col] += (unsigned int)(expr)
#else
#define SG_MMA_CLK(v)
#define SG_MMA_ADD(col, expr)
#endif
// mbarrier wait of the tile pipeline: a protocol error must end the launch (trap -> the host sees a launch failure), not
// hang the cooperative grid
__device__ __forceinline__ void mma_wait(unsigned long long* bar, unsigned int parity) {
    unsigned int done;
    unsigned int spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem columns dcol .. dcol+N) (Mi rows) (+)= A . B^T over K, 3xTF32.  A is a master operand; B a master operand or a
// weight image.  All 256 threads call; returns with every MMA complete and visible to tcgen05.ld.
__device__ __forceinline__ void mma_gemm(MmaSmem& S, MmaPipe& P, int ring_bytes, int MRP, uint32_t tbase, uint32_t dcol, int Mi, int N, int K,
                                      MmaOperand A, MmaOperand B, uint32_t accum) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a_mn = A.kind & 1, b_mn = B.kind & 1;
    const bool b_img = B.kind >= 2;
    const bool a_dir = A.kind == 0;                                  // hi pass straight from the master
    const int pk = A.E * (a_dir ? 4 : 8) + B.E * 8;                  // stage bytes per k
    // chunk length: aim at 4 (weights by TMA: their L2 latency has to be covered) or 3 stages, within what the fillers'
    // plan holds per chunk
    int KC = (ring_bytes / ((b_img ? 4 : 3) * pk)) & ~7;
    if (KC < 8) KC = 8;
    { const int cap = (kMmaMaxG * kMmaFillThreads * 4 / A.E) & ~7; if (KC > cap) KC = cap; }
    if (!b_img) { const int cap = (kMmaMaxG * kMmaFillThreads * 4 / B.E) & ~7; if (KC > cap) KC = cap; }
    if (KC > K) KC = K;
    const int nchunks = (K + KC - 1) / KC;
    int NS = ring_bytes / (KC * pk);
    if (NS > kMmaMaxStages) NS = kMmaMaxStages;
    if (NS > nchunks) NS = nchunks;
    const int stage_floats = (KC * pk) >> 2;
    const int a_img_floats = a_dir ? 1 : 2;                          // x A.E x kc: floats of A's images in a stage
    const unsigned int done_base = P.done_base, fill_base = P.fill_base, full_base = P.full_base;

    SG_MMA_CLK(g0);
    if (warp > 0) {
        // ---- fillers (their first thread also requests the weight chunks: a stage is refilled by whoever sees it free, so
        //      the MMA issuer never waits for a completion) ------------------------------------------------------------------
        const int ftid = tid - 32;
        auto issue_tma = [&](int cn) {
            const int sn = cn % NS;
            const int kn0 = cn * KC, kcn = K - kn0 < KC ? K - kn0 : KC;
            float* Bn = S.stage0 + sn * stage_floats + a_img_floats * A.E * kcn;
            const unsigned int bytes = (unsigned int)(kcn * B.E * 4);
            mbar_expect_tx(S.full + sn, 2u * bytes);
            tma_bulk_g2s(Bn, B.img + (size_t)kn0 * B.E, bytes, S.full + sn);
            tma_bulk_g2s(Bn + B.E * kcn, B.img + B.img_lo + (size_t)kn0 * B.E, bytes, S.full + sn);
        };
        if (b_img && ftid == 0) {
            fence_proxy_async();
            for (int c = 0; c < NS; ++c) issue_tma(c);
        }
        MmaPlan pa, pb;
        mma_plan(A, MRP, ftid, pa);
        if (!b_img) mma_plan(B, MRP, ftid, pb);
        int s = 0, u = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int k0 = c * KC, kc = K - k0 < KC ? K - k0 : KC;
            SG_MMA_CLK(f0);
            if (u > 0) {
                mma_wait(S.done + s, ((done_base >> s) ^ (unsigned int)(u - 1)) & 1u);      // chunk c - NS has been read
                if (b_img && ftid == 0) issue_tma(c);
            }
            SG_MMA_CLK(f1);
            float* st = S.stage0 + s * stage_floats;
            float* Alo = a_dir ? st : st + A.E * kc;
            float* Bhi = Alo + A.E * kc;
            mma_fill(A, pa, MRP, ftid, k0, kc, st, Alo);
            if (!b_img) mma_fill(B, pb, MRP, ftid, k0, kc, Bhi, Bhi + B.E * kc);
            mma::fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(S.filled + s);
#ifdef SG_MMA_PROFILE
            if (tid == 32) { SG_MMA_CLK(f2); SG_MMA_ADD(4, f1 - f0); SG_MMA_ADD(5, f2 - f1); }
#endif
            if (++s == NS) { s = 0; ++u; }
        }
    } else {
        // ---- control warp: one thread issues the weight copies and the MMAs --------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = mma::make_idesc_tf32(Mi, N, a_mn, b_mn);
            // descriptor templates (everything but the start address) and the per-MMA (K = 8) address step in 16-byte units
            const uint64_t a_tmpl = mma::make_desc(0u, a_mn ? 512u : 16u * A.E, a_mn ? 16u * A.E : 128u, a_mn);
            const uint64_t b_tmpl = mma::make_desc(0u, b_mn ? 512u : 16u * B.E, b_mn ? 16u * B.E : 128u, b_mn);
            const uint64_t m_tmpl = mma::make_desc(0u, 16u * MRP, 128u, 0u);                  // the master as a K-major operand
            const uint32_t a_step = 2u * A.E, b_step = 2u * B.E, m_step = 2u * MRP;
            const uint32_t m_base = a_dir ? mma::smem_addr(A.m4) >> 4 : 0u;
            int s = 0, u = 0;
            SG_MMA_ADD(8, clock64() - g0);
            for (int c = 0; c < nchunks; ++c) {
                const int k0 = c * KC, kc = K - k0 < KC ? K - k0 : KC;
                SG_MMA_CLK(i0);
                mma_wait(S.filled + s, ((fill_base >> s) ^ (unsigned int)u) & 1u);
                SG_MMA_CLK(i1);
                if (b_img) mma_wait(S.full + s, ((full_base >> s) ^ (unsigned int)u) & 1u);
                SG_MMA_CLK(i2);
                mma::fence_after_sync();
                float* st = S.stage0 + s * stage_floats;
                const uint32_t al = mma::smem_addr(a_dir ? st : st + A.E * kc) >> 4;
                const uint32_t ah = a_dir ? m_base + (uint32_t)(k0 >> 2) * (uint32_t)MRP : mma::smem_addr(st) >> 4;
                const uint32_t bh = mma::smem_addr(st + a_img_floats * A.E * kc) >> 4, bl = bh + (uint32_t)((B.E * kc) >> 2);
                const uint64_t ah_tmpl = a_dir ? m_tmpl : a_tmpl;
                const uint32_t ah_step = a_dir ? m_step : a_step;
                const int nks = kc >> 3;
#pragma unroll 2
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t dAh = ah_tmpl | (uint64_t)((ah + ks * ah_step) & 0x3FFFu), dAl = a_tmpl | (uint64_t)((al + ks * a_step) & 0x3FFFu);
                    const uint64_t dBh = b_tmpl | (uint64_t)((bh + ks * b_step) & 0x3FFFu), dBl = b_tmpl | (uint64_t)((bl + ks * b_step) & 0x3FFFu);
                    mma::mma_tf32(tbase + dcol, dAl, dBh, idesc, accum);
                    mma::mma_tf32(tbase + dcol, dAh, dBl, idesc, 1u);
                    mma::mma_tf32(tbase + dcol, dAh, dBh, idesc, 1u);
                    accum = 1u;
                }
                mma::commit(S.done + s);
                SG_MMA_CLK(i3);
#ifdef SG_MMA_PROFILE
                SG_MMA_ADD(0, i1 - i0); SG_MMA_ADD(1, i2 - i1); SG_MMA_ADD(2, i3 - i2);
#endif
                if (++s == NS) { s = 0; ++u; }
            }
#ifdef SG_MMA_PROFILE
            SG_MMA_ADD(6, nchunks); sg_mma_prof[P.gid][7] = (unsigned int)(NS * 1000 + KC);
#endif
        }
        __syncwarp();
    }
    // every MMA issued above has completed once the last chunk's commit has arrived
    {
        const int cl = nchunks - 1, sl = cl % NS, ul = cl / NS;
        SG_MMA_CLK(d0);
        mma_wait(S.done + sl, ((done_base >> sl) ^ (unsigned int)ul) & 1u);
        mma::fence_after_sync();
#ifdef SG_MMA_PROFILE
        if (tid == 0) { SG_MMA_ADD(9, clock64() - d0); }
#endif
    }
    // advance the phase bases by the number of times each stage was used (NS <= nchunks: every stage was)
    for (int s = 0; s < NS; ++s) {
        const unsigned int flip = (unsigned int)(((nchunks - 1 - s) / NS + 1) & 1) << s;
        P.done_base ^= flip;
        P.fill_base ^= flip;
        if (b_img) P.full_base ^= flip;
    }
#ifdef SG_MMA_PROFILE
    if (P.gid < 19) ++P.gid;
#endif
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
    mma::fence_async_smem();            // a master written here may be the next contraction's in-place hi operand
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

// tanh(x) = 1 - 2 / (exp(2x) + 1): ex2.approx + rcp.approx, absolute error ~1e-7 (the 1e-4 loss contract has room for it;
// tanhf costs ~5x the instructions and these epilogues sit on the serial path of a job)
__device__ __forceinline__ float mma_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

// weight-gradient epilogues: accumulator row = hidden unit u0 + m -------------------------------------------------------------
// transposed head gradient: G[k*H + u] for k < NA
__device__ __forceinline__ void mma_store_dwh(uint32_t tbase, uint32_t dcol, int Mi, int u0, float* __restrict__ gW, int H, int NA, bool acc) {
    mma_epilogue(tbase, dcol, Mi, 32, [&](int m, int c0, const float (&v)[16]) {
        const int u = u0 + m;
        float old[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) old[j] = (acc && c0 + j < NA) ? __ldcg(gW + (size_t)(c0 + j) * H + u) : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c0 + j < NA) __stcg(gW + (size_t)(c0 + j) * H + u, v[j] + old[j]);
    });
}
__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// G[(u0+m)*ld + c] for c < ncols, N accumulator columns (N % 32 == 0).  The accumulator comes out of tensor memory one ROW per
// thread, but a row of G is what is contiguous in memory: written straight from there, every warp store would touch 32
// different lines.  Each warp therefore turns its 32 x 32 block around in a private staging tile of the (idle) stage ring
// (row pitch 36 floats: conflict-free for the float4 writes and for the float4 / scalar reads) and writes whole 128-byte row
// segments.  acc: add to what an earlier job of this CTA stored -- fire-and-forget reductions (the slot has a single writer,
// so the sum stays deterministic) instead of a load-add-store round trip through L2.  Ends with a CTA barrier.
constexpr int kMmaStorePitch = 36;
__device__ __forceinline__ void mma_store_dw(float* __restrict__ ring, uint32_t tbase, uint32_t dcol, int Mi, int N, int u0,
                                             float* __restrict__ gW, int ld, int ncols, bool acc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int nrows = Mi == 128 ? 32 : 16;              // accumulator rows held by this warp's lanes (M = 64: lanes 0-15)
    const int rbase = nrows * q;
    float* st = ring + warp * (32 * kMmaStorePitch);
    const bool vec = (ld & 3) == 0;
    for (int c0 = 32 * half; c0 < N; c0 += 64) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            float v[16];
            mma::tmem_ld16(tbase + ((uint32_t)(32 * q) << 16) + dcol + (uint32_t)(c0 + 16 * h2), v);
            mma::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(st + lane * kMmaStorePitch + 16 * h2 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
        if (vec) {
            const int cq = lane & 7, c = c0 + 4 * cq;
            if (c < ncols) {
                for (int r = lane >> 3; r < nrows; r += 4) {
                    const float4 o = *reinterpret_cast<const float4*>(st + r * kMmaStorePitch + 4 * cq);
                    float* p = gW + (size_t)(u0 + rbase + r) * ld + c;
                    if (acc) red_add4(p, o);
                    else __stcg(reinterpret_cast<float4*>(p), o);
                }
            }
        } else {
            const int c = c0 + lane;
            if (c < ncols) {
                for (int r = 0; r < nrows; ++r) {
                    const float o = st[r * kMmaStorePitch + lane];
                    float* p = gW + (size_t)(u0 + rbase + r) * ld + c;
                    if (acc) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(o) : "memory");
                    else __stcg(p, o);
                }
            }
        }
        __syncwarp();
    }
    mma::fence_before_sync();
    mma::fence_async_smem();            // the staging tiles were generic writes into the stage ring
    __syncthreads();
}

#ifdef SG_MMA_PROFILE
#define SG_MMA_LAP() do { if (nstage < 16) tstage[nstage++] = clock64(); } while (0)
#else
#define SG_MMA_LAP()
#endif

// Pull the observation (and action) rows of a later job of this CTA towards L2 while the current one computes: the gather at
// the head of a job is a burst of dependent, randomly placed HBM reads with nothing to overlap them with.
template <int MR>
__device__ __forceinline__ void ppo_mma_prefetch_rows(const PpoArgs& a, int step, int tile, int net) {
    const int tid = threadIdx.x;
    if (tid >= MR) return;
    const int epoch = step / a.nmb, mb = step - epoch * a.nmb;
    const int row = a.row_begin + tile * MR + tid;
    if (row >= a.row_end) return;
    const int i = (a.perm + (size_t)epoch * a.S + (size_t)mb * a.mbs)[row];
    const char* p = reinterpret_cast<const char*>(a.obs + (size_t)i * a.O);
    const int bytes = a.O * 4;
    for (int off = 0; off < bytes; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + bytes - 4));
    if (net == 0) {
        const char* q = reinterpret_cast<const char*>(a.actions + (size_t)i * a.A);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + a.A * 4 - 4));
    }
}

// a plain global load the compiler cannot speculate above its guard (asm volatile)
__device__ __forceinline__ float ld_guarded_f32(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- one job ------------------------------------------------------------------------------------------------------------------
// first: no earlier job of this CTA in this optimizer step (gradient accumulators start from zero)
template <int MR>
__device__ void ppo_mma_job(const PpoArgs& a, const MmaDims& d, MmaSmem& S, MmaPipe& P, uint32_t tbase, int step, int tile,
                            int net, float* __restrict__ gout, bool first) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int O = a.O, H = a.H, A = a.A, MRP = d.MRP, RB = d.ring;
    const int epoch = step / a.nmb, mb = step - epoch * a.nmb;
    const int32_t* idx = a.perm + (size_t)epoch * a.S + (size_t)mb * a.mbs;
    const int row0 = a.row_begin + tile * MR;
    const PolicyLayout& L = a.L;
    const int NA = net ? 1 : A;
    const int NA16 = round_up(NA, 16), NA8 = round_up(NA, 8);
    const MmaImg g = make_mma_img(O, H);
    const float* img = a.wimg + (size_t)(2 * net) * g.net_floats;
    const int lo = g.net_floats;
    const bool acc = !first;
#ifdef SG_MMA_PROFILE
    long long tstage[16];
    int nstage = 0;
    const bool prof = blockIdx.x == 0 && tid == 0 && step == 1;
    P.gid = 0;
    __syncthreads();
    for (int i = tid; i < 20 * 10; i += kStepThreads) (&sg_mma_prof[0][0])[i] = 0u;
    __syncthreads();
#endif
    SG_MMA_LAP();
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
    // ---- gather observation (and action) rows into the masters.  A warp takes whole rows: its lanes read consecutive floats of
    //      one sampled row (whole 128-byte lines whatever obs_dim is) and scatter them into the [col/4][row][col%4] master --
    //      with the odd row pitch the 32 scalar stores of an instruction fall into 32 different banks.  The head of a job is a
    //      burst of dependent, randomly placed HBM reads with nothing to overlap them with: what counts is how many rows are
    //      in flight per warp (measured at obs_dim 111, 128 rows: 4 rows 11.5k cycles, 8 rows 7.5k; 4-byte cp.async copies of
    //      everything at once 10k -- their issue rate binds; touch loads / a bulk L2 prefetch of the next job's rows: no gain).
    {
        float* Xf = reinterpret_cast<float*>(S.X);
        const int ncx = d.Op8;                                  // master columns incl. zero padding
        constexpr int GB = 8, NW = kStepThreads / 32;
        for (int cb = 0; cb < ncx; cb += 128) {
            for (int r0 = warp; r0 < MR; r0 += GB * NW) {
                float v[GB][4];
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    const int r = r0 + u * NW;
                    const int i = r < MR ? S.IDX[r] : -1;
                    const float* src = a.obs + (size_t)(i < 0 ? 0 : i) * O + cb + lane;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[u][j] = 0.f;
                        if (i >= 0 && cb + lane + 32 * j < O) v[u][j] = ld_guarded_f32(src + 32 * j);
                    }
                }
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    const int r = r0 + u * NW;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = cb + lane + 32 * j;
                        if (r < MR && c < ncx) Xf[(((c >> 2) * MRP + r) << 2) + (c & 3)] = v[u][j];
                    }
                }
            }
        }
        if (net == 0) {
            float* Af = reinterpret_cast<float*>(S.ACT);
            const int nca = 4 * d.ag;                           // <= 32
            for (int r0 = warp; r0 < MR; r0 += GB * NW) {
                float v[GB];
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    const int r = r0 + u * NW;
                    const int i = r < MR ? S.IDX[r] : -1;
                    v[u] = 0.f;
                    if (i >= 0 && lane < A) v[u] = ld_guarded_f32(a.actions + (size_t)i * A + lane);
                }
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    const int r = r0 + u * NW;
                    if (r < MR && lane < nca) Af[(((lane >> 2) * MRP + r) << 2) + (lane & 3)] = v[u];
                }
            }
        }
    }
    mma::fence_async_smem();            // the X master is G1's in-place hi operand
    __syncthreads();
    SG_MMA_LAP();   // 1: gather

    // ---- forward (A2C/model.py:255-264) ---------------------------------------------------------------------------------------
    mma_gemm(S, P, RB, MRP, tbase, 0, MR, H, d.Op8, op_master_k(S.X, MR), op_image(img + g.w1k, lo, H, 0), 0u);
    SG_MMA_LAP();   // 2: G1
    mma_epilogue(tbase, 0, MR, H, [&](int r, int c0, const float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            S.H1[(c >> 2) * MRP + r] = make_float4(mma_tanh(v[4 * j] + S.B1[c]), mma_tanh(v[4 * j + 1] + S.B1[c + 1]),
                                                   mma_tanh(v[4 * j + 2] + S.B1[c + 2]), mma_tanh(v[4 * j + 3] + S.B1[c + 3]));
        }
    });
    SG_MMA_LAP();   // 3: E1
    mma_gemm(S, P, RB, MRP, tbase, 0, MR, H, H, op_master_k(S.H1, MR), op_image(img + g.w2k, lo, H, 0), 0u);
    SG_MMA_LAP();   // 4: G2
    mma_epilogue(tbase, 0, MR, H, [&](int r, int c0, const float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            S.H2[(c >> 2) * MRP + r] = make_float4(mma_tanh(v[4 * j] + S.B2[c]), mma_tanh(v[4 * j + 1] + S.B2[c + 1]),
                                                   mma_tanh(v[4 * j + 2] + S.B2[c + 2]), mma_tanh(v[4 * j + 3] + S.B2[c + 3]));
        }
    });
    SG_MMA_LAP();   // 5: E2
    // head: Gaussian mean (A2C/distributions.py:109-110) or critic_linear
    mma_gemm(S, P, RB, MRP, tbase, 0, MR, NA16, H, op_master_k(S.H2, MR), op_image(img + g.whk, lo, NA16, 0), 0u);
    SG_MMA_LAP();   // 6: G3

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
                // log-prob of the stored action, summed over the action dim (A2C/distributions.py:52-53); v[k] <- a - mu.
                // The constant part sum_k(-log sigma_k - log sqrt(2 pi)) sits in LSG[31]; 1/(2 var), 1/var are per-step tables.
                float lp = S.LSG[31];
#pragma unroll
                for (int kq = 0; kq < 8; ++kq) {
                    if (4 * kq < A) {
                        const float4 av = live ? S.ACT[kq * MRP + r] : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float ac[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = 4 * kq + j;
                            if (k < A) {
                                const float dd = ac[j] - (v[k] + S.BH[k]);
                                v[k] = dd;
                                lp = fmaf(-dd * dd, S.I2VAR[k], lp);
                            }
                        }
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
                // d logp / d mu = (a-mu)/var -> dHead; d logp / d logstd = (a-mu)^2/var - 1 -> scratch rows (summed over the tile's
                // rows below by all warps)
                float4* DL = reinterpret_cast<float4*>(S.stage0);
#pragma unroll
                for (int kq = 0; kq < 8; ++kq) {
                    float dl[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = 4 * kq + j;
                        const float dd = v[k], iv = k < A ? S.IVAR[k] : 0.f;
                        v[k] = (ok && k < A) ? coef * dd * iv : 0.f;
                        dl[j] = (ok && k < A) ? coef * (dd * dd * iv - 1.f) : 0.f;
                    }
                    if (live && 4 * kq < A) DL[kq * MRP + r] = make_float4(dl[0], dl[1], dl[2], dl[3]);
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
                        const float gg = l1 > l2 ? e1 : (l2 > l1 ? in2 * e2 : 0.5f * (e1 + in2 * e2));
                        dv = a.c_v * invB * gg;
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
        if (tid >= 32 && tid < 34) {
            const float s = S.RED[tid] + S.RED[kMmaRedStride + tid] + S.RED[2 * kMmaRedStride + tid] + S.RED[3 * kMmaRedStride + tid];
            S.LOSS[tid - 32] += s;
        }
        if (net == 0) mma_colsum(reinterpret_cast<const float4*>(S.stage0), MRP, MR, d.ag, S.GLS);
        mma_colsum(S.DH, MRP, MR, net == 0 ? d.ag : 1, S.GBH);
        mma::fence_async_smem();            // stage 0 was generic scratch; its next writer may be a TMA copy
    }
    SG_MMA_LAP();   // 7: losses

    // ---- backward ------------------------------------------------------------------------------------------------------------------
    const uint32_t accum = (d.tacc && !first) ? 1u : 0u;
    // head weight gradient, transposed: D(unit, a) = sum_rows H2(row, unit) dHead(row, a)
    for (int b = 0; b < d.nblk; ++b) {
        mma_gemm(S, P, RB, MRP, tbase, d.tacc ? d.c_dwh : 0, d.Mb, 32, MR, op_master_mn(S.H2, d.Mb, b * d.Mb, H), op_master_mn(S.DH, 32, 0, 32), accum);
        if (!d.tacc) mma_store_dwh(tbase, 0, d.Mb, b * d.Mb, gout + (net ? L.vw : L.mw), H, NA, acc);
    }
    SG_MMA_LAP();   // 8: G4 + E4
    // dZ2 = (dHead . Wh) * (1 - h2^2), in place over the H2 master
    mma_gemm(S, P, RB, MRP, tbase, 0, MR, H, NA8, op_master_k(S.DH, MR), op_image(img + g.whm, lo, H, 1), 0u);
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
    SG_MMA_LAP();   // 9: G5 + E5
    // dW2(n, k) = sum_rows dZ2(row, n) H1(row, k); with two unit blocks the N extent is halved as well (stage capacity)
    {
        const int nsplit = (d.nblk > 1 && 2 * 8 * 8 * (d.Mb + H) > d.ring) ? 2 : 1, Nh = H / nsplit;      // two stages of K = 8 must fit
        for (int b = 0; b < d.nblk; ++b)
            for (int h = 0; h < nsplit; ++h) {
                mma_gemm(S, P, RB, MRP, tbase, d.tacc ? d.c_dw2 : 0, d.Mb, Nh, MR, op_master_mn(S.H2, d.Mb, b * d.Mb, H),
                         op_master_mn(S.H1, Nh, h * Nh, H), accum);
                if (!d.tacc) mma_store_dw(S.stage0, tbase, 0, d.Mb, Nh, b * d.Mb, gout + (net ? L.cw2 : L.aw2) + h * Nh, H, Nh, acc);
            }
    }
    SG_MMA_LAP();   // 10: G6 + E6
    // dZ1 = (dZ2 . W2) * (1 - h1^2), in place over the H1 master
    mma_gemm(S, P, RB, MRP, tbase, 0, MR, H, H, op_master_k(S.H2, MR), op_image(img + g.w2m, lo, H, 1), 0u);
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
    SG_MMA_LAP();   // 11: G7 + E7
    // dW1(n, k) = sum_rows dZ1(row, n) X(row, k)
    for (int b = 0; b < d.nblk; ++b) {
        mma_gemm(S, P, RB, MRP, tbase, d.tacc ? d.c_dw1 : 0, d.Mb, d.Op32, MR, op_master_mn(S.H1, d.Mb, b * d.Mb, H),
                 op_master_mn(S.X, d.Op32, 0, d.Op8), accum);
        if (!d.tacc) mma_store_dw(S.stage0, tbase, 0, d.Mb, d.Op32, b * d.Mb, gout + (net ? L.cw1 : L.aw1), O, O, acc);
    }
    SG_MMA_LAP();   // 12: G8 + E8
#ifdef SG_MMA_PROFILE
    if (prof) {
        printf("job tile %d net %d first %d cycles:", tile, net, (int)first);
        for (int i = 1; i < nstage; ++i) printf(" %lld", tstage[i] - tstage[i - 1]);
        printf("\n");
        for (int g = 0; g < P.gid; ++g)
            printf("  gemm %2d: chunks %3u NS*1000+KC %5u | issuer: setup %5u wait_filled %6u wait_full %6u issue %6u drain %5u | filler: wait_done %6u fill %6u\n",
                   g, sg_mma_prof[g][6], sg_mma_prof[g][7], sg_mma_prof[g][8], sg_mma_prof[g][0], sg_mma_prof[g][1], sg_mma_prof[g][2],
                   sg_mma_prof[g][9], sg_mma_prof[g][4], sg_mma_prof[g][5]);
    }
#endif
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
        const float ls = tid < A ? ld_cg(a.params + L.ls + tid) : 0.f;
        const float sigma = expf(ls), var = sigma * sigma;
        S.LS[tid] = ls; S.IVAR[tid] = 1.f / var; S.I2VAR[tid] = 1.f / (2.f * var);
        // LSG[31] = sum_k (-log sigma_k - log sqrt(2 pi)) in the order the per-element form adds them
        float c = tid < A ? -logf(sigma) - SG_LOG_SQRT_2PI : 0.f;
        c = warp_sum(c);
        S.LSG[tid] = c;
        S.GBH[tid] = 0.f; S.GLS[tid] = 0.f;
        if (tid < 4) S.LOSS[tid] = 0.f;
    }
    __syncthreads();
    float* gout = a.gpart + (size_t)cta * a.P;
    bool first = true;
    for (int job = cta; job < njobs; job += ncta) {
        if (job + ncta < njobs) ppo_mma_prefetch_rows<MR>(a, step, (job + ncta) >> 1, net);
        else if (step + 1 < a.nsteps) ppo_mma_prefetch_rows<MR>(a, step + 1, cta >> 1, net);
        ppo_mma_job<MR>(a, d, S, P, tbase, step, job >> 1, net, gout, first);
        first = false;
    }
    if (d.tacc) {
        // the CTA's weight gradients of this step, accumulated in tensor memory over its jobs
        mma_store_dwh(tbase, d.c_dwh, d.Mb, 0, gout + (net ? L.vw : L.mw), H, NA, false);
        mma_store_dw(S.stage0, tbase, d.c_dw2, d.Mb, H, 0, gout + (net ? L.cw2 : L.aw2), H, H, false);
        mma_store_dw(S.stage0, tbase, d.c_dw1, d.Mb, d.Op32, 0, gout + (net ? L.cw1 : L.aw1), a.O, a.O, false);
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
