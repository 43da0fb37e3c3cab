// Host side of the minibatch sampler: torch.randperm(S) streams of the CPU default generator, bit for bit, off the critical path.
//
// The reference draws one torch.randperm(T*N) per PPO epoch from the CPU default generator (A2C/storage.py:158-162 through
// BatchSampler(SubsetRandomSampler(range(batch_size)))).  For n < 2^32 / 20 ATen's randperm_cpu is a Fisher-Yates walk fed by
// one 32-bit draw of the generator's mt19937 engine per element ("z = random() % (n - i); swap(r[i], r[i + z])").  At the
// synthetic 1024x4096 rollout that is a 4.2 M-step chain of dependent cache misses, ~40-110 ms per epoch on one core -- longer
// than the epoch's kernel -- and the reference contract ("bit-exact rollout indexing/sampling given identical seeds") rules out
// a different sampler.  This file reproduces the SAME stream faster and asynchronously:
//   * one generator thread runs the mt19937 engine (state handed in by the caller from torch.get_rng_state()) and emits the
//     n - 1 draws of every requested permutation, in order -- the only inherently sequential part (~2 ns per draw);
//   * worker threads turn each draw block into its permutation: the moduli are computed a few iterations ahead and the swap
//     partner r[i + z] is prefetched, so the walk runs at cache bandwidth instead of cache latency, and the walks of different
//     epochs run concurrently;
//   * the caller waits per permutation (the kernel of epoch e starts as soon as permutation e exists) and gets the final
//     engine state back to store into the torch generator, which then is exactly where torch.randperm would have left it.
// Host-only code (no kernel); tests/test_host_logic.py pins it against torch.randperm including the generator state.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "sg_common.cuh"

namespace sg {

// mt19937 state transition and tempering, written so that the host compiler vectorises them: within a pass, word k reads the
// OLD words k+1 and k+397 (first 227 words) or the NEW word k-227 (the rest) -- the dependence distance is far beyond any vector
// width.  Two copies: baseline x86-64 and, picked at run time, AVX2 (8 words per instruction).
#define SG_MT_TWIST(dst, a, b, far) do { const uint32_t y_ = ((a) & 0x80000000u) | ((b) & 0x7fffffffu); \
                                         (dst) = (far) ^ (y_ >> 1) ^ ((uint32_t)(-(int32_t)(y_ & 1u)) & 0x9908b0dfu); } while (0)
#define SG_MT_REGENERATE_BODY                                                            \
    for (int k = 0; k < 624 - 397; ++k) SG_MT_TWIST(key[k], key[k], key[k + 1], key[k + 397]);        \
    for (int k = 624 - 397; k < 623; ++k) SG_MT_TWIST(key[k], key[k], key[k + 1], key[k + 397 - 624]); \
    SG_MT_TWIST(key[623], key[623], key[0], key[396]);
#define SG_MT_TEMPER_BODY                                                                \
    for (int64_t j = 0; j < count; ++j) {                                                \
        uint32_t y = in[j];                                                              \
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18; \
        out[j] = y;                                                                      \
    }
static void mt_regenerate_base(uint32_t* __restrict__ key) { SG_MT_REGENERATE_BODY }
static void mt_temper_base(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int64_t count) { SG_MT_TEMPER_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void mt_regenerate_avx2(uint32_t* __restrict__ key) { SG_MT_REGENERATE_BODY }
__attribute__((target("avx2"))) static void mt_temper_avx2(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int64_t count) { SG_MT_TEMPER_BODY }
static bool mt_have_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
#else
static void mt_regenerate_avx2(uint32_t* key) { mt_regenerate_base(key); }
static void mt_temper_avx2(uint32_t* out, const uint32_t* in, int64_t count) { mt_temper_base(out, in, count); }
static bool mt_have_avx2() { return false; }
#endif

struct Mt19937 {
    uint32_t key[624];
    int pos;                       // next word to temper; 624 = regenerate first
    void regenerate() {
        if (mt_have_avx2()) mt_regenerate_avx2(key);
        else mt_regenerate_base(key);
        pos = 0;
    }
    // advance by count draws without producing them
    void discard(int64_t count) {
        while (count > 0) {
            if (pos >= 624) regenerate();
            int64_t take = 624 - pos;
            if (take > count) take = count;
            pos += (int)take;
            count -= take;
        }
    }
    void fill(uint32_t* out, int64_t count) {
        int64_t i = 0;
        while (i < count) {
            if (pos >= 624) regenerate();
            int64_t take = 624 - pos;
            if (take > count - i) take = count - i;
            if (mt_have_avx2()) mt_temper_avx2(out + i, key + pos, take);
            else mt_temper_base(out + i, key + pos, take);
            pos += (int)take;
            i += take;
        }
    }
};

// r = the permutation ATen's randperm_cpu builds from these n - 1 draws (draws is scratch: overwritten with the moduli)
static void fisher_yates(int32_t* r, uint32_t* draws, int64_t n) {
    for (int64_t i = 0; i < n; ++i) r[i] = (int32_t)i;
    constexpr int64_t D = 24;                 // look-ahead of the modulus + prefetch
    const int64_t m = n - 1;
    for (int64_t i = 0; i < (D < m ? D : m); ++i) {
        draws[i] = draws[i] % (uint32_t)(n - i);
        __builtin_prefetch(r + i + draws[i], 1);
    }
    for (int64_t i = 0; i < m; ++i) {
        const int64_t a = i + D;
        if (a < m) {
            draws[a] = draws[a] % (uint32_t)(n - a);
            __builtin_prefetch(r + a + draws[a], 1);
        }
        const int64_t j = i + draws[i];
        const int32_t sav = r[i];
        r[i] = r[j];
        r[j] = sav;
    }
}

struct RandpermJob {
    Mt19937 mt;
    int64_t n = 0;
    int n_perms = 0;
    int32_t* out = nullptr;
    uint64_t owned = ~0ull;                   // bit e: permutation e is built here (others only move the engine)
    std::vector<std::vector<uint32_t>> draws;
    std::vector<int> drawn, done;             // guarded by m
    int next = 0;                             // next permutation a worker takes (guarded by m)
    std::mutex m;
    std::condition_variable cv;
    std::thread gen;
    std::vector<std::thread> workers;

    void generator() {
        for (int e = 0; e < n_perms; ++e) {
            std::vector<uint32_t> d;
            const bool mine = (owned >> e) & 1ull;
            if (mine) {
                d.resize((size_t)(n > 1 ? n - 1 : 0));
                mt.fill(d.data(), (int64_t)d.size());
            } else {
                mt.discard(n > 1 ? n - 1 : 0);
            }
            {
                std::lock_guard<std::mutex> lk(m);
                draws[e] = std::move(d);
                drawn[e] = 1;
                if (!mine) done[e] = 1;
            }
            cv.notify_all();
        }
    }
    void worker() {
        for (;;) {
            int e;
            std::vector<uint32_t> d;
            {
                std::unique_lock<std::mutex> lk(m);
                if (next >= n_perms) return;
                e = next++;
                cv.wait(lk, [&] { return drawn[e] != 0; });
                if (!((owned >> e) & 1ull)) continue;
                d = std::move(draws[e]);
            }
            fisher_yates(out + (size_t)e * n, d.data(), n);
            {
                std::lock_guard<std::mutex> lk(m);
                done[e] = 1;
            }
            cv.notify_all();
        }
    }
};

}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

void* sg_host_randperm_begin(const uint32_t* mt_key, int mt_pos, int64_t n, int n_perms, int32_t* out, int n_threads,
                             uint64_t owned_mask) {
    if (!mt_key || !out || n < 1 || n_perms < 1 || n_perms > 64 || mt_pos < 0 || mt_pos > 624 || n >= (int64_t)(0xffffffffu / 20u)) {
        set_error("sg_host_randperm_begin: bad arguments (n in [1, 2^32/20): above that ATen switches algorithm; at most 64 permutations per stream)");
        return nullptr;
    }
    RandpermJob* j = new RandpermJob();
    std::memcpy(j->mt.key, mt_key, sizeof(j->mt.key));
    j->mt.pos = mt_pos;
    j->n = n; j->n_perms = n_perms; j->out = out; j->owned = owned_mask;
    j->draws.resize(n_perms);
    j->drawn.assign(n_perms, 0);
    j->done.assign(n_perms, 0);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_perms) n_threads = n_perms;
    j->gen = std::thread([j] { j->generator(); });
    for (int t = 0; t < n_threads; ++t) j->workers.emplace_back([j] { j->worker(); });
    return j;
}

int sg_host_randperm_prefix(const uint32_t* mt_key, int mt_pos, int64_t n, int64_t m, int32_t* out, uint32_t* mt_key_out,
                            int* mt_pos_out) {
    SG_REQUIRE(mt_key && out && n >= 1 && m >= 0 && m <= n && mt_pos >= 0 && mt_pos <= 624 && n < (int64_t)(0xffffffffu / 20u),
               "sg_host_randperm_prefix: bad arguments");
    Mt19937 mt;
    std::memcpy(mt.key, mt_key, sizeof(mt.key));
    mt.pos = mt_pos;
    // element i of the result is final after step i of the walk, and step i only reads positions >= i: the first m elements
    // need the first min(m, n-1) steps; the remaining draws only move the engine
    const int64_t steps = m < n - 1 ? m : n - 1;
    std::vector<uint32_t> d((size_t)steps);
    mt.fill(d.data(), steps);
    mt.discard(n - 1 - steps);
    if (steps > n / 8) {
        // a long prefix: the plain dense walk
        std::vector<int32_t> r((size_t)n);
        for (int64_t i = 0; i < n; ++i) r[i] = (int32_t)i;
        for (int64_t i = 0; i < steps; ++i) {
            const int64_t j = i + d[i] % (uint32_t)(n - i);
            const int32_t sav = r[i];
            r[i] = r[j];
            r[j] = sav;
        }
        std::memcpy(out, r.data(), (size_t)m * sizeof(int32_t));
        if (mt_key_out) std::memcpy(mt_key_out, mt.key, sizeof(mt.key));
        if (mt_pos_out) *mt_pos_out = mt.pos;
        return SG_OK;
    }
    // the walk only ever touches positions i < steps and the `steps` partners i + z: positions below m live in `head`, the
    // others in a small open-addressing table (absent = still the identity), so nothing of size n is allocated
    std::vector<int32_t> head((size_t)m);
    for (int64_t i = 0; i < m; ++i) head[i] = (int32_t)i;
    size_t cap = 64;
    while (cap < (size_t)steps * 2 + 2) cap <<= 1;
    std::vector<int64_t> tkey(cap, -1);
    std::vector<int32_t> tval(cap);
    auto slot = [&](int64_t pos) -> size_t {
        size_t h = ((uint64_t)pos * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
        while (tkey[h] != -1 && tkey[h] != pos) h = (h + 1) & (cap - 1);
        return h;
    };
    for (int64_t i = 0; i < steps; ++i) {
        const int64_t j = i + d[i] % (uint32_t)(n - i);
        if (j < m) {
            const int32_t sav = head[i];
            head[i] = head[j];
            head[j] = sav;
        } else {
            const size_t h = slot(j);
            const int32_t vj = tkey[h] == j ? tval[h] : (int32_t)j;
            tkey[h] = j;
            tval[h] = head[i];
            head[i] = vj;
        }
    }
    std::memcpy(out, head.data(), (size_t)m * sizeof(int32_t));
    if (mt_key_out) std::memcpy(mt_key_out, mt.key, sizeof(mt.key));
    if (mt_pos_out) *mt_pos_out = mt.pos;
    return SG_OK;
}

int sg_host_randperm_wait(void* handle, int e) {
    RandpermJob* j = (RandpermJob*)handle;
    SG_REQUIRE(j && e >= 0 && e < j->n_perms, "sg_host_randperm_wait: bad handle or index");
    std::unique_lock<std::mutex> lk(j->m);
    j->cv.wait(lk, [&] { return j->done[e] != 0; });
    return SG_OK;
}

int sg_host_randperm_end(void* handle, uint32_t* mt_key_out, int* mt_pos_out) {
    RandpermJob* j = (RandpermJob*)handle;
    SG_REQUIRE(j, "sg_host_randperm_end: null handle");
    j->gen.join();
    for (auto& w : j->workers) w.join();
    if (mt_key_out) std::memcpy(mt_key_out, j->mt.key, sizeof(j->mt.key));
    if (mt_pos_out) *mt_pos_out = j->mt.pos;
    delete j;
    return SG_OK;
}

#pragma GCC visibility pop
}
