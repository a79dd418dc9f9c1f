// tables.cpp -- see tables.h.  Host only (no CUDA): compiled by nvcc's host compiler into the same library.
#include "tables.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>

namespace holdem {
namespace {

// Strength of a five-card class as one integer: category in the top digit, then up to five tie-break ranks,
// most significant first (base 16).  Larger = stronger.  Categories: 8 straight flush, 7 quads, 6 full house,
// 5 flush, 4 straight, 3 trips, 2 two pair, 1 pair, 0 high card.
uint32_t straight_top(uint32_t mask) {  // 0 if the 5-bit mask is not a straight, else 1 + rank of its top card
    if (mask == 0x100Fu) return 1 + 3;  // A-2-3-4-5: five high
    for (int lo = 0; lo + 4 <= 12; ++lo)
        if (mask == (0x1Fu << lo)) return 1 + lo + 4;
    return 0;
}

uint32_t strength5(const int cnt[13], bool suited) {
    // ranks ordered by (multiplicity, rank) descending
    std::array<int, 5> order{};
    int m = 0;
    uint32_t mask = 0;
    for (int c = 4; c >= 1; --c)
        for (int r = 12; r >= 0; --r)
            if (cnt[r] == c) order[m++] = r;
    for (int r = 0; r < 13; ++r)
        if (cnt[r]) mask |= 1u << r;
    int top = cnt[order[0]], second = m > 1 ? cnt[order[1]] : 0;
    uint32_t cat;
    uint32_t tie = 0;
    if (top == 1) {
        uint32_t st = straight_top(mask);
        if (st) {
            cat = suited ? 8 : 4;
            tie = (st - 1) << 16;
            return (cat << 20) | tie;
        }
        cat = suited ? 5 : 0;
    } else if (top == 4) cat = 7;
    else if (top == 3) cat = second == 2 ? 6 : 3;
    else cat = second == 2 ? 2 : 1;
    for (int i = 0; i < m; ++i) tie |= (uint32_t)order[i] << (16 - 4 * i);
    return (cat << 20) | tie;
}

struct Class5 {
    uint32_t strength;
    uint32_t key;   // additive key (unsuited) or 13-bit mask (suited)
    bool suited;
};

bool build_hash(const std::vector<uint32_t> &keys, const std::vector<uint16_t> &vals, int bucket_bits,
                int slot_bits, PerfectHash &ph) {
    const uint32_t nb = 1u << bucket_bits, ns = 1u << slot_bits;
    uint32_t mul = 0x9E3779B1u;
    for (int attempt = 0; attempt < 64; ++attempt) {
        uint32_t mul_b = mul | 1u;
        mul = mul * 0x2545F491u + 0x6B43A9B5u;
        uint32_t mul_s = mul | 1u;
        mul = mul * 0x2545F491u + 0x6B43A9B5u;
        std::vector<std::vector<uint32_t>> buckets(nb);
        for (uint32_t i = 0; i < keys.size(); ++i) buckets[(keys[i] * mul_b) >> (32 - bucket_bits)].push_back(i);
        std::vector<uint32_t> order(nb);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t a, uint32_t b) { return buckets[a].size() > buckets[b].size(); });
        std::vector<uint16_t> val(ns, 0), disp(nb, 0);
        std::vector<uint8_t> used(ns, 0);
        bool ok = true;
        for (uint32_t b : order) {
            const auto &items = buckets[b];
            if (items.empty()) break;
            std::vector<uint32_t> h(items.size());
            for (size_t k = 0; k < items.size(); ++k) h[k] = (keys[items[k]] * mul_s) >> (32 - slot_bits);
            uint32_t d = 0;
            for (; d < ns && d < 65536u; ++d) {
                bool fits = true;
                for (size_t k = 0; k < h.size() && fits; ++k) {
                    uint32_t s = (h[k] + d) & (ns - 1);
                    if (used[s]) fits = false;
                    for (size_t q = 0; q < k && fits; ++q)
                        if (((h[q] + d) & (ns - 1)) == s) fits = false;
                }
                if (fits) break;
            }
            if (d >= ns || d >= 65536u) { ok = false; break; }
            disp[b] = (uint16_t)d;
            for (size_t k = 0; k < h.size(); ++k) {
                uint32_t s = (h[k] + d) & (ns - 1);
                used[s] = 1;
                val[s] = vals[items[k]];
            }
        }
        if (!ok) continue;
        ph.mul_b = mul_b; ph.mul_s = mul_s;
        ph.bucket_bits = bucket_bits; ph.slot_bits = slot_bits;
        ph.disp = std::move(disp); ph.val = std::move(val);
        return true;
    }
    return false;
}

// Work list over m positions: every (a,b,d), a<b<d-1<m-1, grouped by walk length d-b-1 into 32-lane tasks (see tables.h).
void build_walk_list(int m, std::vector<uint32_t> &hdr_out, std::vector<uint32_t> &lanes_out) {
    struct Task { uint32_t hdr; std::vector<uint32_t> lanes; };
    std::vector<Task> tasks;
    for (int L = m - 3; L >= 1; --L) {
        std::vector<uint32_t> group;
        for (int b = m - 2 - L; b >= 1; --b)
            for (int a = 0; a < b; ++a) group.push_back((uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)(b + 1 + L) << 16));
        for (size_t s0 = 0; s0 < group.size(); s0 += 32) {
            std::vector<uint32_t> lanes(32, 0xFFFFFFFFu);
            for (size_t l = 0; l < 32 && s0 + l < group.size(); ++l) lanes[l] = group[s0 + l];
            // the kernels' packed 6-bit counters take at most 31 iterations between drains: split long walks
            for (int k0 = 0; k0 < L; k0 += 31)
                tasks.push_back({(uint32_t)k0 | ((uint32_t)std::min(31, L - k0) << 8), lanes});
        }
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const Task &x, const Task &y) { return (x.hdr >> 8) > (y.hdr >> 8); });
    for (const Task &tk : tasks) {
        hdr_out.push_back(tk.hdr);
        lanes_out.insert(lanes_out.end(), tk.lanes.begin(), tk.lanes.end());
    }
    while (hdr_out.size() % 4) hdr_out.push_back(0);  // 16-byte multiple; iters 0 = no-op
    lanes_out.resize(hdr_out.size() * 32, 0xFFFFFFFFu);
}

struct Builder {
    Tables t;
    bool ok = false;
    char err[256] = {0};

    void run() {
        // ---- 1. the 7462 five-card classes -----------------------------------------------------------
        std::vector<Class5> classes;
        int cnt[13];
        // unsuited rank multisets of size 5, at most 4 of a rank
        std::vector<int> seq(5);
        for (seq[0] = 0; seq[0] < 13; ++seq[0])
            for (seq[1] = seq[0]; seq[1] < 13; ++seq[1])
                for (seq[2] = seq[1]; seq[2] < 13; ++seq[2])
                    for (seq[3] = seq[2]; seq[3] < 13; ++seq[3])
                        for (seq[4] = seq[3]; seq[4] < 13; ++seq[4]) {
                            if (seq[0] == seq[4]) continue;  // five of a rank
                            std::memset(cnt, 0, sizeof cnt);
                            uint32_t key = 0;
                            for (int r : seq) { cnt[r]++; key += kRankKey[r]; }
                            classes.push_back({strength5(cnt, false), key, false});
                        }
        // suited: 5 distinct ranks
        for (uint32_t mask = 0; mask < 8192; ++mask) {
            if (__builtin_popcount(mask) != 5) continue;
            std::memset(cnt, 0, sizeof cnt);
            for (int r = 0; r < 13; ++r)
                if (mask >> r & 1) cnt[r] = 1;
            classes.push_back({strength5(cnt, true), mask, true});
        }
        if ((int)classes.size() != kNumClasses) { std::snprintf(err, sizeof err, "class count %zu", classes.size()); return; }
        std::sort(classes.begin(), classes.end(),
                  [](const Class5 &a, const Class5 &b) { return a.strength > b.strength; });
        for (size_t i = 1; i < classes.size(); ++i)
            if (classes[i].strength == classes[i - 1].strength) { std::snprintf(err, sizeof err, "strength tie"); return; }

        // ---- 2. dense scratch tables keyed directly by the additive key -----------------------------------
        uint32_t max_key[3] = {0, 0, 0};
        for (int n = 5; n <= 7; ++n) {  // largest sum: four aces + (n-4) kings
            max_key[n - 5] = 4 * kRankKey[12] + (n - 4) * kRankKey[11];
        }
        std::vector<uint16_t> dense[3];
        for (int i = 0; i < 3; ++i) dense[i].assign(max_key[i] + 1, 0);
        t.flush.assign(kFlushTableSize, 0);
        for (size_t i = 0; i < classes.size(); ++i) {
            uint16_t rank = (uint16_t)(i + 1);
            if (classes[i].suited) t.flush[classes[i].key] = rank;
            else dense[0][classes[i].key] = rank;
        }
        // flush table for 6..13 suited cards: best five = minimum over dropping one card
        for (uint32_t mask = 0; mask < 8192; ++mask) {  // ascending order: sub-masks are smaller numbers
            if (__builtin_popcount(mask) <= 5) continue;
            uint16_t best = 0xFFFF;
            for (int r = 0; r < 13; ++r)
                if (mask >> r & 1) best = std::min(best, t.flush[mask & ~(1u << r)]);
            t.flush[mask] = best;
        }
        // ---- 3. 6- and 7-card multisets: drop-one-card dynamic programme ---------------------------------------
        std::vector<uint32_t> keys[3];
        std::vector<uint16_t> vals[3];
        for (int n = 5; n <= 7; ++n) {
            std::vector<int> s(n, 0);
            // iterate non-decreasing sequences of length n over 0..12
            while (true) {
                std::memset(cnt, 0, sizeof cnt);
                uint32_t key = 0;
                bool valid = true;
                for (int r : s) { if (++cnt[r] > 4) valid = false; key += kRankKey[r]; }
                if (valid) {
                    uint16_t v;
                    if (n == 5) v = dense[0][key];
                    else {
                        v = 0xFFFF;
                        for (int r = 0; r < 13; ++r)
                            if (cnt[r]) v = std::min(v, dense[n - 6][key - kRankKey[r]]);
                        if (dense[n - 5][key] != 0) { std::snprintf(err, sizeof err, "key collision n=%d", n); return; }
                        dense[n - 5][key] = v;
                    }
                    if (v == 0 || v == 0xFFFF) { std::snprintf(err, sizeof err, "missing class n=%d", n); return; }
                    keys[n - 5].push_back(key);
                    vals[n - 5].push_back(v);
                }
                int i = n - 1;
                while (i >= 0 && s[i] == 12) --i;
                if (i < 0) break;
                int v2 = s[i] + 1;
                for (int j = i; j < n; ++j) s[j] = v2;
            }
            t.n_keys[n - 5] = (int)keys[n - 5].size();
        }
        // distinctness of the additive keys within each size
        for (int i = 0; i < 3; ++i) {
            std::vector<uint32_t> k = keys[i];
            std::sort(k.begin(), k.end());
            if (std::adjacent_find(k.begin(), k.end()) != k.end()) { std::snprintf(err, sizeof err, "duplicate key"); return; }
        }
        // ---- 4. perfect hashes --------------------------------------------------------------------------------
        const int bucket_bits[3] = {11, 13, 14}, slot_bits[3] = {13, 15, 16};
        for (int i = 0; i < 3; ++i)
            if (!build_hash(keys[i], vals[i], bucket_bits[i], slot_bits[i], t.nf[i])) {
                std::snprintf(err, sizeof err, "perfect hash failed n=%d", i + 5);
                return;
            }
        for (int i = 0; i < 3; ++i)
            for (size_t k = 0; k < keys[i].size(); ++k)
                if (t.nf[i].lookup(keys[i][k]) != vals[i][k]) { std::snprintf(err, sizeof err, "hash verify n=%d", i + 5); return; }
        // ---- 5. colex pair table -------------------------------------------------------------------------------
        t.pairs.resize(kMaxPairs);
        int p = 0;
        for (int j = 1; j < 52; ++j)
            for (int i = 0; i < j; ++i) t.pairs[p++] = (uint16_t)(i | (j << 8));
        // ---- 6. flop kernel work list (47 positions) and the reference-mode lists for 46 / 45 positions -----------------
        build_walk_list(47, t.flop_task_hdr, t.flop_lanes);
        build_walk_list(46, t.ref_task_hdr[0], t.ref_lanes[0]);
        build_walk_list(45, t.ref_task_hdr[1], t.ref_lanes[1]);
        // ---- 6b. turn kernel work list -----------------------------------------------------------------------------------
        {
            struct Task { uint32_t hdr; std::vector<uint32_t> lanes; };
            std::vector<Task> tasks;
            for (int b = 1; b <= 44; ++b) {           // walk length 45 - b, b lanes (a = 0..b-1)
                const int L = 45 - b;
                for (int s0 = 0; s0 < b; s0 += 32) {
                    std::vector<uint32_t> lanes(32, 0xFFFFFFFFu);
                    for (int l = 0; l < 32 && s0 + l < b; ++l) lanes[l] = (uint32_t)(s0 + l) | ((uint32_t)b << 8);
                    for (int k0 = 0; k0 < L; k0 += 31)
                        tasks.push_back({(uint32_t)k0 | ((uint32_t)std::min(31, L - k0) << 8), lanes});
                }
            }
            std::stable_sort(tasks.begin(), tasks.end(),
                             [](const Task &x, const Task &y) { return (x.hdr >> 8) > (y.hdr >> 8); });
            for (const Task &tk : tasks) {
                t.turn_task_hdr.push_back(tk.hdr);
                t.turn_lanes.insert(t.turn_lanes.end(), tk.lanes.begin(), tk.lanes.end());
            }
            while (t.turn_task_hdr.size() % 4) t.turn_task_hdr.push_back(0);
            t.turn_lanes.resize(t.turn_task_hdr.size() * 32, 0xFFFFFFFFu);
            t.nf3_ranks.assign(456, 0);
            for (int r0 = 0; r0 < 13; ++r0)
                for (int r1 = r0; r1 < 13; ++r1)
                    for (int r2 = r1; r2 < 13; ++r2)
                        t.nf3_ranks[r0 + (r1 + 1) * r1 / 2 + (r2 + 2) * (r2 + 1) * r2 / 6] =
                            (uint16_t)(r0 | (r1 << 4) | (r2 << 8));
        }
        // ---- 7. the 1820 multisets of four ranks ---------------------------------------------------------------------------
        t.nf4_ranks.assign(1824, 0);
        for (int r0 = 0; r0 < 13; ++r0)
            for (int r1 = r0; r1 < 13; ++r1)
                for (int r2 = r1; r2 < 13; ++r2)
                    for (int r3 = r2; r3 < 13; ++r3) {
                        int e = r0 + (r1 + 1) * r1 / 2 + (r2 + 2) * (r2 + 1) * r2 / 6 +
                                (r3 + 3) * (r3 + 2) * (r3 + 1) * r3 / 24;
                        t.nf4_ranks[e] = (uint16_t)(r0 | (r1 << 4) | (r2 << 8) | (r3 << 12));
                    }
        // ---- 9. direct-indexed board / own-hand tables of the rank-multiset flop kernel ------------------------------------------
        {
            auto rank_of = [&](const int *ranks, int n, int hash) -> uint16_t {   // 0 when a rank would be needed five times
                int cnt[13] = {0};
                uint32_t key = 0;
                for (int i = 0; i < n; ++i) {
                    if (++cnt[ranks[i]] > 4) return 0;
                    key += kRankKey[ranks[i]];
                }
                return (uint16_t)t.nf[hash].lookup(key);
            };
            int px[91], py[91];
            for (int y = 0; y < 13; ++y)
                for (int x = 0; x <= y; ++x) { px[y * (y + 1) / 2 + x] = x; py[y * (y + 1) / 2 + x] = y; }
            t.board_t.assign((size_t)kBoardMultisets * kBoardTRow, 0);
            t.board_opp5.assign((size_t)kBoardMultisets * 96, 0);
            for (int r2 = 0; r2 < 13; ++r2)
                for (int r1 = 0; r1 <= r2; ++r1)
                    for (int r0 = 0; r0 <= r1; ++r0) {
                        const int b3 = r0 + (r1 + 1) * r1 / 2 + (r2 + 2) * (r2 + 1) * r2 / 6;
                        uint16_t *row = t.board_t.data() + (size_t)b3 * kBoardTRow;
                        for (int q = 0; q < 91; ++q) {
                            for (int p = q; p < 91; ++p) {
                                const int r[7] = {r0, r1, r2, px[q], py[q], px[p], py[p]};
                                const uint16_t v = rank_of(r, 7, 2);
                                row[q * kBoardTPitch + p] = v;
                                row[p * kBoardTPitch + q] = v;
                            }
                            const int r5[5] = {r0, r1, r2, px[q], py[q]};
                            t.board_opp5[(size_t)b3 * 96 + q] = rank_of(r5, 5, 0);
                        }
                    }
            t.own_our7.assign((size_t)kOwnMultisets * 96, 0);
            t.own_rank5.assign(kOwnMultisets + 4, 0);
            for (int r4 = 0; r4 < 13; ++r4)
                for (int r3 = 0; r3 <= r4; ++r3)
                    for (int r2 = 0; r2 <= r3; ++r2)
                        for (int r1 = 0; r1 <= r2; ++r1)
                            for (int r0 = 0; r0 <= r1; ++r0) {
                                const int m5 = r0 + (r1 + 1) * r1 / 2 + (r2 + 2) * (r2 + 1) * r2 / 6 +
                                               (r3 + 3) * (r3 + 2) * (r3 + 1) * r3 / 24 +
                                               (r4 + 4) * (r4 + 3) * (r4 + 2) * (r4 + 1) * r4 / 120;
                                const int r5[5] = {r0, r1, r2, r3, r4};
                                t.own_rank5[m5] = rank_of(r5, 5, 0);
                                for (int q = 0; q < 91; ++q) {
                                    const int r[7] = {r0, r1, r2, r3, r4, px[q], py[q]};
                                    t.own_our7[(size_t)m5 * 96 + q] = rank_of(r, 7, 2);
                                }
                            }
            // ---- four-card boards (turn) ---------------------------------------------------------------------------------
            t.board4_t.assign((size_t)kBoard4Multisets * kBoard4TRow, 0);
            t.board4_opp6.assign((size_t)kBoard4Multisets * 96, 0);
            for (int r3 = 0; r3 < 13; ++r3)
                for (int r2 = 0; r2 <= r3; ++r2)
                    for (int r1 = 0; r1 <= r2; ++r1)
                        for (int r0 = 0; r0 <= r1; ++r0) {
                            const int b4 = r0 + (r1 + 1) * r1 / 2 + (r2 + 2) * (r2 + 1) * r2 / 6 +
                                           (r3 + 3) * (r3 + 2) * (r3 + 1) * r3 / 24;
                            for (int p = 0; p < 91; ++p) {
                                for (int rq = 0; rq < 13; ++rq) {
                                    const int r[7] = {r0, r1, r2, r3, rq, px[p], py[p]};
                                    t.board4_t[(size_t)b4 * kBoard4TRow + p * 16 + rq] = rank_of(r, 7, 2);
                                }
                                const int r6[6] = {r0, r1, r2, r3, px[p], py[p]};
                                t.board4_opp6[(size_t)b4 * 96 + p] = rank_of(r6, 6, 1);
                            }
                        }
            t.own6_our7.assign((size_t)kOwn6Multisets * 16, 0);
            t.own6_rank6.assign(kOwn6Multisets + 4, 0);
            for (int r5 = 0; r5 < 13; ++r5)
                for (int r4 = 0; r4 <= r5; ++r4)
                    for (int r3 = 0; r3 <= r4; ++r3)
                        for (int r2 = 0; r2 <= r3; ++r2)
                            for (int r1 = 0; r1 <= r2; ++r1)
                                for (int r0 = 0; r0 <= r1; ++r0) {
                                    const int m6 = r0 + (r1 + 1) * r1 / 2 + (r2 + 2) * (r2 + 1) * r2 / 6 +
                                                   (r3 + 3) * (r3 + 2) * (r3 + 1) * r3 / 24 +
                                                   (r4 + 4) * (r4 + 3) * (r4 + 2) * (r4 + 1) * r4 / 120 +
                                                   (r5 + 5) * (r5 + 4) * (r5 + 3) * (r5 + 2) * (r5 + 1) * r5 / 720;
                                    const int r6[6] = {r0, r1, r2, r3, r4, r5};
                                    t.own6_rank6[m6] = rank_of(r6, 6, 1);
                                    for (int rq = 0; rq < 13; ++rq) {
                                        const int r[7] = {r0, r1, r2, r3, r4, r5, rq};
                                        t.own6_our7[(size_t)m6 * 16 + rq] = rank_of(r, 7, 2);
                                    }
                                }
        }
        ok = true;
    }
};

Builder &builder() {
    static Builder b;
    static std::once_flag once;
    std::call_once(once, [] { b.run(); });
    return b;
}

uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

}  // namespace

const Tables &get_tables() { return builder().t; }

int selftest(char *msg, int msg_len) {
    Builder &b = builder();
    auto fail = [&](const char *m) {
        if (msg && msg_len > 0) std::snprintf(msg, msg_len, "%s", m);
        return 1;
    };
    if (!b.ok) return fail(b.err[0] ? b.err : "table build failed");
    const Tables &t = b.t;
    if (t.n_keys[0] != 6175 || t.n_keys[1] != 18395 || t.n_keys[2] != 49205) return fail("multiset counts");
    // every rank 1..7462 appears exactly once among the 5-card flush masks and the 5-card unsuited multisets
    std::vector<int> seen(kNumClasses + 1, 0);
    int n_flush5 = 0;
    for (uint32_t m = 0; m < 8192; ++m)
        if (__builtin_popcount(m) == 5) { seen[t.flush[m]]++; n_flush5++; }
    if (n_flush5 != 1287) return fail("flush count");
    for (uint16_t v : t.nf[0].val)
        if (v) seen[v]++;
    for (int r = 1; r <= kNumClasses; ++r)
        if (seen[r] != 1) return fail("classes are not a bijection onto 1..7462");
    // category boundaries (sizes 10/156/156/1277/10/858/858/2860/1277)
    if (t.flush[0x1F00] != 1 || t.flush[0x100F] != 10 || t.flush[0x1F00 >> 1] != 2) return fail("straight flush ranks");
    if (t.flush[0x1E80] != 323 || t.flush[0x002F] != 1599) return fail("flush boundaries");
    uint32_t k_quads_top = 4 * kRankKey[12] + kRankKey[11];   // AAAA K
    uint32_t k_worst = kRankKey[5] + kRankKey[3] + kRankKey[2] + kRankKey[1] + kRankKey[0];  // 7 5 4 3 2
    uint32_t k_broadway = kRankKey[12] + kRankKey[11] + kRankKey[10] + kRankKey[9] + kRankKey[8];
    if (t.nf[0].lookup(k_quads_top) != 11) return fail("quads boundary");
    if (t.nf[0].lookup(k_worst) != 7462) return fail("worst hand");
    if (t.nf[0].lookup(k_broadway) != 1600) return fail("broadway");
    if (t.pairs[0] != (0 | 1 << 8) || t.pairs[1] != (0 | 2 << 8) || t.pairs[2] != (1 | 2 << 8)) return fail("pairs");
    {   // the flop work list covers every 4-set of 47 positions exactly once
        uint64_t sets = 0;
        for (size_t tk = 0; tk < t.flop_task_hdr.size(); ++tk) {
            uint32_t k0 = t.flop_task_hdr[tk] & 0xFF, iters = t.flop_task_hdr[tk] >> 8;
            if (iters > 31) return fail("task too long");
            for (int l = 0; l < 32; ++l) {
                uint32_t v = t.flop_lanes[tk * 32 + l];
                if (v == 0xFFFFFFFFu) continue;
                uint32_t a = v & 0xFF, b = (v >> 8) & 0xFF, d = (v >> 16) & 0xFF;
                if (!(a < b && b + 1 + k0 + iters <= d && d <= 46)) return fail("bad flop lane");
                sets += iters;
            }
        }
        if (sets != 178365) return fail("flop work list does not cover C(47,4)");
        for (int q = 0; q < 2; ++q) {
            sets = 0;
            for (size_t tk = 0; tk < t.ref_task_hdr[q].size(); ++tk)
                for (int l = 0; l < 32; ++l)
                    if (t.ref_lanes[q][tk * 32 + l] != 0xFFFFFFFFu) sets += t.ref_task_hdr[q][tk] >> 8;
            if (sets != (q == 0 ? 163185u : 148995u)) return fail("reference work list does not cover C(m,4)");
        }
        sets = 0;
        for (size_t tk = 0; tk < t.turn_task_hdr.size(); ++tk) {
            uint32_t k0 = t.turn_task_hdr[tk] & 0xFF, iters = t.turn_task_hdr[tk] >> 8;
            for (int l = 0; l < 32; ++l) {
                uint32_t v = t.turn_lanes[tk * 32 + l];
                if (v == 0xFFFFFFFFu) continue;
                uint32_t a = v & 0xFF, b = (v >> 8) & 0xFF;
                if (!(a < b && b + k0 + iters <= 45 && iters <= 31)) return fail("bad turn lane");
                sets += iters;
            }
        }
        if (sets != 15180) return fail("turn work list does not cover C(46,3)");
        for (int e = 0; e < 1820; ++e) {
            uint16_t v = t.nf4_ranks[e];
            int r0 = v & 15, r1 = (v >> 4) & 15, r2 = (v >> 8) & 15, r3 = v >> 12;
            if (!(r0 <= r1 && r1 <= r2 && r2 <= r3 && r3 < 13)) return fail("nf4 ranks");
        }
    }
    {   // direct-indexed tables: spot values against the hashes they were built from
        if (t.board_t.size() != (size_t)kBoardMultisets * kBoardTRow || t.own_our7.size() != (size_t)kOwnMultisets * 96)
            return fail("board table sizes");
        // board 2 3 4 (ranks 0,1,2 -> b3 = 0 + 1 + 4), q = pair (5,6) = ranks (3,4), p = (A,A): straight 2-6, rank 1608
        const int b3 = 0 + (1 + 1) * 1 / 2 + (2 + 2) * (2 + 1) * 2 / 6;
        const int q = 4 * 5 / 2 + 3, p = 12 * 13 / 2 + 12;
        const uint32_t key = kRankKey[0] + kRankKey[1] + kRankKey[2] + kRankKey[3] + kRankKey[4] + 2 * kRankKey[12];
        if (t.board_t[(size_t)b3 * kBoardTRow + q * kBoardTPitch + p] != t.nf[2].lookup(key) ||
            t.board_t[(size_t)b3 * kBoardTRow + p * kBoardTPitch + q] != t.nf[2].lookup(key))
            return fail("board_t entry");
        if (t.nf[2].lookup(key) != 1608) return fail("six-high straight");
        // four deuces on the board side are impossible with a pair of deuces added twice
        if (t.board_t[0 * kBoardTRow + 0 * kBoardTPitch + 0] != 0) return fail("impossible multiset must be 0");
        // own hand Q K K A A (ranks 10,11,11,12,12) and the same hand plus the rank pair (Q,Q)
        const int r[5] = {10, 11, 11, 12, 12};
        const int m5 = r[0] + (r[1] + 1) * r[1] / 2 + (r[2] + 2) * (r[2] + 1) * r[2] / 6 +
                       (r[3] + 3) * (r[3] + 2) * (r[3] + 1) * r[3] / 24 + (r[4] + 4) * (r[4] + 3) * (r[4] + 2) * (r[4] + 1) * r[4] / 120;
        const uint32_t k5 = kRankKey[10] + 2 * kRankKey[11] + 2 * kRankKey[12];
        if (t.own_rank5[m5] != t.nf[0].lookup(k5)) return fail("own_rank5 entry");
        const int qq = 10 * 11 / 2 + 10;
        if (t.own_our7[(size_t)m5 * 96 + qq] != t.nf[2].lookup(k5 + 2 * kRankKey[10])) return fail("own_our7 entry");
        // turn tables: board 2 3 4 5 (b4 = 0 + 1 + 4 + 15), river 6 (rank 4), holding (A,A): the six-high straight again
        const int b4 = 0 + (1 + 1) * 1 / 2 + (2 + 2) * (2 + 1) * 2 / 6 + (3 + 3) * (3 + 2) * (3 + 1) * 3 / 24;
        if (t.board4_t.size() != (size_t)kBoard4Multisets * kBoard4TRow || t.own6_our7.size() != (size_t)kOwn6Multisets * 16)
            return fail("board4 table sizes");
        if (t.board4_t[(size_t)b4 * kBoard4TRow + p * 16 + 4] != 1608) return fail("board4_t entry");
        if (t.board4_opp6[(size_t)b4 * 96 + p] != t.nf[1].lookup(kRankKey[0] + kRankKey[1] + kRankKey[2] + kRankKey[3] + 2 * kRankKey[12]))
            return fail("board4_opp6 entry");
    }
    if (msg && msg_len > 0) msg[0] = 0;
    return 0;
}

TableImage make_image(const Tables &t) {
    TableImage im;
    auto append = [&](const void *src, size_t n) {
        uint32_t off = align16((uint32_t)im.bytes.size());
        im.bytes.resize(off + n);
        std::memcpy(im.bytes.data() + off, src, n);
        return off;
    };
    im.off_flush = append(t.flush.data(), t.flush.size() * 2);
    im.off_pairs = append(t.pairs.data(), t.pairs.size() * 2);
    im.off_task_hdr = append(t.flop_task_hdr.data(), t.flop_task_hdr.size() * 4);
    im.off_lanes = append(t.flop_lanes.data(), t.flop_lanes.size() * 4);
    im.off_nf4 = append(t.nf4_ranks.data(), t.nf4_ranks.size() * 2);
    im.off_board_t = append(t.board_t.data(), t.board_t.size() * 2);
    im.off_board_opp5 = append(t.board_opp5.data(), t.board_opp5.size() * 2);
    im.off_own_our7 = append(t.own_our7.data(), t.own_our7.size() * 2);
    im.off_own_rank5 = append(t.own_rank5.data(), t.own_rank5.size() * 2);
    im.off_board4_t = append(t.board4_t.data(), t.board4_t.size() * 2);
    im.off_board4_opp6 = append(t.board4_opp6.data(), t.board4_opp6.size() * 2);
    im.off_own6_our7 = append(t.own6_our7.data(), t.own6_our7.size() * 2);
    im.off_own6_rank6 = append(t.own6_rank6.data(), t.own6_rank6.size() * 2);
    im.n_tasks = (uint32_t)t.flop_task_hdr.size();
    im.off_turn_hdr = append(t.turn_task_hdr.data(), t.turn_task_hdr.size() * 4);
    im.off_turn_lanes = append(t.turn_lanes.data(), t.turn_lanes.size() * 4);
    im.off_nf3 = append(t.nf3_ranks.data(), t.nf3_ranks.size() * 2);
    im.n_turn_tasks = (uint32_t)t.turn_task_hdr.size();
    for (int q = 0; q < 2; ++q) {
        im.off_ref_hdr[q] = append(t.ref_task_hdr[q].data(), t.ref_task_hdr[q].size() * 4);
        im.off_ref_lanes[q] = append(t.ref_lanes[q].data(), t.ref_lanes[q].size() * 4);
        im.n_ref_tasks[q] = (uint32_t)t.ref_task_hdr[q].size();
    }
    for (int i = 0; i < 3; ++i) {
        im.off_disp[i] = append(t.nf[i].disp.data(), t.nf[i].disp.size() * 2);
        im.off_val[i] = append(t.nf[i].val.data(), t.nf[i].val.size() * 2);
    }
    im.bytes.resize(align16((uint32_t)im.bytes.size()));
    return im;
}

}  // namespace holdem
