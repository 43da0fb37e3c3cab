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

struct Mt19937 {
    uint32_t key[624];
    int pos;                       // next word to temper; 624 = regenerate first
    void regenerate() {
        constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
        int k = 0;
        for (; k < 624 - 397; ++k) {
            const uint32_t y = (key[k] & kUpper) | (key[k + 1] & kLower);
            key[k] = key[k + 397] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        }
        for (; k < 623; ++k) {
            const uint32_t y = (key[k] & kUpper) | (key[k + 1] & kLower);
            key[k] = key[k + 397 - 624] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        }
        const uint32_t y = (key[623] & kUpper) | (key[0] & kLower);
        key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        pos = 0;
    }
    static inline uint32_t temper(uint32_t y) {
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
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
            for (int64_t j = 0; j < take; ++j) out[i + j] = temper(key[pos + j]);
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
    std::vector<std::vector<uint32_t>> draws;
    std::vector<int> drawn, done;             // guarded by m
    int next = 0;                             // next permutation a worker takes (guarded by m)
    std::mutex m;
    std::condition_variable cv;
    std::thread gen;
    std::vector<std::thread> workers;

    void generator() {
        for (int e = 0; e < n_perms; ++e) {
            std::vector<uint32_t> d((size_t)(n > 1 ? n - 1 : 0));
            mt.fill(d.data(), (int64_t)d.size());
            {
                std::lock_guard<std::mutex> lk(m);
                draws[e] = std::move(d);
                drawn[e] = 1;
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

void* sg_host_randperm_begin(const uint32_t* mt_key, int mt_pos, int64_t n, int n_perms, int32_t* out, int n_threads) {
    if (!mt_key || !out || n < 1 || n_perms < 1 || mt_pos < 0 || mt_pos > 624 || n >= (int64_t)(0xffffffffu / 20u)) {
        set_error("sg_host_randperm_begin: bad arguments (n must be in [1, 2^32/20): above that ATen switches algorithm)");
        return nullptr;
    }
    RandpermJob* j = new RandpermJob();
    std::memcpy(j->mt.key, mt_key, sizeof(j->mt.key));
    j->mt.pos = mt_pos;
    j->n = n; j->n_perms = n_perms; j->out = out;
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
