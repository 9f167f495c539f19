// seir_prologue_host.cuh -- the reference's prologue picks for one seed, on the host (seir_individual.py:163, :176-187).
//
// The reference seeds numba's MT19937 and draws, per non-empty age group, np.random.choice(matches, k, replace=False)
// and finally np.random.choice(n, k0, replace=False).  numba implements both as a Fisher-Yates shuffle of a copy
// (numba/cpython/randomimpl.py:1927-1946, 2084-2100: i = len-1 .. 1, j = randint(0, i) by masked rejection of 32-bit
// words, swap) and returns its first k entries -- two full passes over n draws whatever k is, because the stream has to
// advance.  The drop-in run_model must hand the SAME agents to the device as the reference would pick for that seed.
//
// What this restatement does not do is shuffle: the first k entries of the shuffled array are found by following k
// POSITIONS backwards through the swaps.  Undoing the swaps in reverse order of execution (i = 1, 2, .., len-1), the
// element that ends at position p sits before swap i at position s_i(q), s_i exchanging i and j_i; it starts at q = p and
// is the identity's element q after the last undo.  So one pass draws the j_i (sequential 4 B stores), a second pass reads them
// in ascending order and tests "is j_i one of the k tracked positions" against a bitmap of len bits (cache resident);
// a hit happens about k (1 + ln(len / k)) times in total.  With k = 0 -- every age group of a run without shelter in
// place -- nothing is stored at all, the draws are only consumed.  A group with k > len / 4 is shuffled for real (the
// groups are n / 101 agents: cache resident).  At n = 10 M this is 0.03 s per seed against 0.3 s for the two
// NumPy RandomState permutations it replaces (tables.prologue, which stays as the checked alternative).
#pragma once
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace seir_host {

struct Mt19937 {   // numba/_random.c:40-75 (init_genrand / genrand_int32 of Matsumoto & Nishimura)
    uint32_t mt[624];
    uint32_t out[624];   // the tempered words of the current block
    int index;
    explicit Mt19937(uint32_t seed)
    {
        for (int pos = 0; pos < 624; pos++) {
            mt[pos] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
        }
        index = 624;
    }
    void refill()
    {
        const uint32_t UP = 0x80000000u, LOW = 0x7fffffffu, A = 0x9908b0dfu;
        int i;
        for (i = 0; i < 624 - 397; i++) {
            const uint32_t y = (mt[i] & UP) | (mt[i + 1] & LOW);
            mt[i] = mt[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        for (; i < 623; i++) {
            const uint32_t y = (mt[i] & UP) | (mt[i + 1] & LOW);
            mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        const uint32_t y = (mt[623] & UP) | (mt[0] & LOW);
        mt[623] = mt[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        for (i = 0; i < 624; i++) {   // tempering of the whole block (vectorised by the compiler)
            uint32_t v = mt[i];
            v ^= v >> 11;
            v ^= (v << 7) & 0x9d2c5680u;
            v ^= (v << 15) & 0xefc60000u;
            v ^= v >> 18;
            out[i] = v;
        }
        index = 0;
    }
    // The shuffle's draws j_i = np.random.randint(0, i + 1) for i = hi, hi-1, .., 1 (randomimpl.py:150-195, 454-521:
    // 32-bit words masked to the bit length of i, rejected while above i; i < 2^32 here).  STORE: keep them in J[i].
    // Written without a data-dependent branch -- every word is masked, stored and accepted by arithmetic -- because the
    // rejection (28 % of the words on average) is unpredictable and a mispredicted branch costs more than the draw.
    template <bool STORE>
    void shuffle_draws(int64_t hi, uint32_t *J)
    {
        int64_t i = hi;
        while (i > 0) {
            // all i in (lo, i] share the mask of i's bit length
            uint32_t mask = (uint32_t)i;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            const int64_t lo = (int64_t)(mask >> 1);   // the largest i with a shorter bit length (0 ends the loop)
            while (i > lo) {
                if (index >= 624) refill();
                int idx = index;
#if defined(__SSE2__)
                // counting only: 32 words at a time.  Over a batch i falls by at most 32, so a word is accepted whatever
                // happened before it if it is <= i - 32 and rejected if it is > i; a word in between (probability 32 / i)
                // sends the batch to the scalar loop below.  (signed compares: the operands stay below 2^31)
                if (!STORE && mask < 0x80000000u) {
                    static const unsigned char bits4[16] = {0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
                    const __m128i vmask = _mm_set1_epi32((int)mask);
                    while (idx + 32 <= 624 && i - 32 > lo) {
                        const __m128i vhi = _mm_set1_epi32((int)i), vlo = _mm_set1_epi32((int)(i - 32));
                        int accepted = 0, unsure = 0;
                        for (int b = 0; b < 8; ++b) {
                            const __m128i r = _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i *>(out + idx + 4 * b)), vmask);
                            const __m128i above_lo = _mm_cmpgt_epi32(r, vlo), above_hi = _mm_cmpgt_epi32(r, vhi);
                            accepted += bits4[15 ^ _mm_movemask_ps(_mm_castsi128_ps(above_lo))];
                            unsure |= _mm_movemask_ps(_mm_castsi128_ps(_mm_andnot_si128(above_hi, above_lo)));
                        }
                        if (unsure) break;
                        i -= accepted;
                        idx += 32;
                    }
                }
#endif
                while (idx < 624 && i > lo) {
                    const uint32_t r = out[idx++] & mask;
                    if (STORE) J[i] = r;
                    i -= (int64_t)(r <= (uint32_t)i);
                }
                index = idx;
            }
        }
    }
};

// The first k entries of the shuffled identity of length len, written to out[0..k).  `J` and `bits` are scratch.
inline void first_k_of_shuffle(Mt19937 &g, int64_t len, int64_t k, int64_t *out, std::vector<uint32_t> &J,
                               std::vector<uint64_t> &bits)
{
    if (len <= 1) {                      // the shuffle loop does not run (i = len-1 .. 1)
        if (k > 0 && len == 1) out[0] = 0;
        return;
    }
    if (k <= 0) {                        // consume the draws, keep nothing
        g.shuffle_draws<false>(len - 1, nullptr);
        return;
    }
    if (k > len / 4) {                   // a real shuffle
        if (J.size() < (size_t)len) J.resize((size_t)len);   // (never shrunk: growing a vector zero-fills)
        g.shuffle_draws<true>(len - 1, J.data());
        std::vector<uint32_t> x((size_t)len);
        for (int64_t i = 0; i < len; i++) x[(size_t)i] = (uint32_t)i;
        for (int64_t i = len - 1; i > 0; i--) {
            const uint32_t j = J[(size_t)i], tmp = x[(size_t)i];
            x[(size_t)i] = x[j];
            x[j] = tmp;
        }
        for (int64_t r = 0; r < k; r++) out[r] = (int64_t)x[(size_t)r];
        return;
    }
    if (J.size() < (size_t)len) J.resize((size_t)len);
    g.shuffle_draws<true>(len - 1, J.data());
    bits.assign((size_t)((len + 63) / 64), 0);
    std::unordered_map<uint32_t, int64_t> slot_at;   // tracked position -> which output it becomes
    slot_at.reserve((size_t)k * 2);
    auto set_bit = [&](uint32_t q) { bits[q >> 6] |= 1ull << (q & 63); };
    auto clr_bit = [&](uint32_t q) { bits[q >> 6] &= ~(1ull << (q & 63)); };
    auto has_bit = [&](uint32_t q) { return (bits[q >> 6] >> (q & 63)) & 1ull; };
    // output 0 is never position i of a swap (i >= 1): it is tracked from the start; output p >= 1 enters at its own
    // undo step i = p, where it still sits at position p
    slot_at[0] = 0;
    set_bit(0);
    for (int64_t i = 1; i < len; i++) {
        const uint32_t j = J[(size_t)i];
        const bool a = i < k;                 // the tracked position equal to i, if any, is output i itself
        const bool b = has_bit(j) != 0;
        if (!a && !b) continue;
        if (j == (uint32_t)i) {               // a swap with itself
            if (a) { slot_at[(uint32_t)i] = i; set_bit((uint32_t)i); }
            continue;
        }
        // exchange what sits at i and at j
        int64_t at_j = -1;
        if (b) {
            auto it = slot_at.find(j);
            at_j = it->second;
            slot_at.erase(it);
            clr_bit(j);
        }
        if (a) { slot_at[j] = i; set_bit(j); }                      // output i moves to j
        if (at_j >= 0) { slot_at[(uint32_t)i] = at_j; set_bit((uint32_t)i); }   // the one from j moves to i
    }
    for (const auto &kv : slot_at) out[kv.second] = (int64_t)kv.first;
}

}  // namespace seir_host

extern "C" int seir_prologue_mt(uint32_t seed, int64_t n, int n_ages, const int64_t *group_offsets,
                                const int64_t *group_members, const double *fraction_stay_home,
                                double initial_infected_fraction, uint8_t *home_real, int64_t *initial_infected,
                                int64_t initial_cap, int64_t *n_initial)
{
    if (n < 0 || n >= (1ll << 32) || n_ages < 0 || !group_offsets || (n > 0 && !group_members) || !home_real || !n_initial)
        return fail("seir_prologue_mt: bad argument");
    const int64_t k0 = (int64_t)(initial_infected_fraction * (double)n);   // :187 int(initial_infected_fraction*n)
    *n_initial = k0;
    if (k0 > 0 && (!initial_infected || initial_cap < k0)) return fail("seir_prologue_mt: initial_infected is too short");
    if (k0 > n) return fail("seir_prologue_mt: more initial cases than agents");   // np.random.choice raises
    try {
        seir_host::Mt19937 g(seed);                                           // :163
        // scratch kept between calls (40 MB at n = 10 M: its page faults cost more than the draws that fill it)
        static thread_local std::vector<uint32_t> J;
        static thread_local std::vector<uint64_t> bits;
        std::vector<int64_t> picks;
        std::memset(home_real, 0, (size_t)n);
        for (int a = 0; a < n_ages; a++) {                                     // :178-182
            const int64_t lo = group_offsets[a], m = group_offsets[a + 1] - lo;
            if (m <= 0) continue;
            if (!fraction_stay_home) { seir_host::first_k_of_shuffle(g, m, 0, nullptr, J, bits); continue; }
            const int64_t k = (int64_t)(fraction_stay_home[a] * (double)m);
            if (k > m) return fail("seir_prologue_mt: fraction_stay_home above 1");
            picks.resize((size_t)(k > 0 ? k : 0));
            seir_host::first_k_of_shuffle(g, m, k, picks.data(), J, bits);
            for (int64_t r = 0; r < k; r++) home_real[group_members[lo + picks[(size_t)r]]] = 1;
        }
        seir_host::first_k_of_shuffle(g, n, k0, initial_infected, J, bits);    // :187
    } catch (const std::exception &e) {
        return fail(e.what());
    }
    return 0;
}
