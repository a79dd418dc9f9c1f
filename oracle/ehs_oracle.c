/*
 * oracle/ehs_oracle.c -- CPU restatement of the reference's hand-strength / hand-potential path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it; the product never does.
 *
 * Two rank semantics:
 *   mode 1 "reference": the reference exactly as written -- CalculateHandRankValue
 *       (HandCalculations.py:45-82, 9 categories, higher = better, multiset semantics, quirks kept),
 *       opponents from deck \ (our u board) (:148-150), runouts from deck \ board ONLY (:152-154).
 *       PINNED against the unmodified reference (tests/golden/literal_*.json).
 *   mode 0 "treys": the canonical Effective-Hand-Strength enumeration the north star names, same loop
 *       structure as HandCalculations.py:157-226 but rank = Treys evaluate (lower = better) and
 *       runouts drawn from the cards unseen by both players.  PARITY UNPINNED by the reference
 *       (it never runs this variant); pinned by treys_oracle.c's known-answer tests only.
 *
 * Integer counts only; the float formulas (HandCalculations.py:175, :221-224) live in oracle/oracle.py
 * so both sides derive floats from counts with the same dtype and operation order.
 */
#include <stdint.h>
#include <string.h>

int treys_evaluate(const int *cards, int n);
void treys_init(void);

/* ---- HandCalculations.py:24-37 is_straight(values) -------------------------------------------- */
int lit_is_straight(const int *values, int n) {
    /* :26-27  any duplicate -> False */
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (values[i] == values[j]) return 0;
    int ext[32];
    /* :29 sorted(values) */
    for (int i = 0; i < n; i++) ext[i] = values[i];
    for (int i = 1; i < n; i++) {
        int v = ext[i], j = i - 1;
        while (j >= 0 && ext[j] > v) { ext[j + 1] = ext[j]; j--; }
        ext[j + 1] = v;
    }
    /* :31 extended = sorted + [v + 13 ...] */
    for (int i = 0; i < n; i++) ext[n + i] = ext[i] + 13;
    /* :33-35 any window of five consecutive list elements that are consecutive integers */
    for (int i = 0; i + 4 < 2 * n; i++) {
        int ok = 1;
        for (int k = 1; k < 5; k++)
            if (ext[i + k] != ext[i] + k) ok = 0;
        if (ok) return 1;
    }
    return 0;
}

/* ---- HandCalculations.py:45-82 CalculateHandRankValue on the concatenated card list ------------- */
int lit_category(const int *cards, int n) {
    int values[16], suits[16];
    int vcnt[13] = {0}, scnt[4] = {0};
    if (n > 16) return -1;
    for (int i = 0; i < n; i++) {
        values[i] = cards[i] % 13; /* :53 */
        suits[i] = cards[i] / 13;  /* :54 */
        vcnt[values[i]]++;         /* :57 */
        scnt[suits[i]]++;          /* :56 */
    }
    int maxc = 0; /* :59 */
    for (int v = 0; v < 13; v++)
        if (vcnt[v] > maxc) maxc = vcnt[v];
    /* :39-43,:61 determine_flush: first suit (Counter insertion order) with count >= 5 */
    int flush = 0, fsuit = -1;
    for (int i = 0; i < n && !flush; i++)
        if (scnt[suits[i]] >= 5) { flush = 1; fsuit = suits[i]; }
    if (flush) { /* :62-65 */
        int fv[16], m = 0;
        for (int i = 0; i < n; i++)
            if (suits[i] == fsuit) fv[m++] = values[i];
        if (lit_is_straight(fv, m)) return 8;
    }
    if (maxc == 4) return 7; /* :68 */
    int has2 = 0, n2 = 0;
    for (int v = 0; v < 13; v++)
        if (vcnt[v] == 2) { has2 = 1; n2++; }
    if (maxc == 3 && has2) return 6;         /* :70 */
    if (lit_is_straight(values, n)) return 4; /* :72 */
    if (flush) return 5;                      /* :74 */
    if (maxc == 3) return 3;                  /* :76 */
    if (maxc == 2 && n2 == 2) return 2;       /* :78 */
    if (maxc == 2) return 1;                  /* :80 */
    return 0;                                 /* :82 */
}

/* ---- HandCalculations.py:131-139 remove_cards_from_deck ------------------------------------------ */
static int deck_without(const int *rm, int nrm, int *deck) {
    int present[52];
    for (int i = 0; i < 52; i++) present[i] = 1;
    for (int i = 0; i < nrm; i++)
        if (rm[i] >= 0 && rm[i] < 52) present[rm[i]] = 0; /* unknown cards are skipped (:135-136) */
    int m = 0;
    for (int i = 0; i < 52; i++)
        if (present[i]) deck[m++] = i;
    return m;
}

static int rank_of(int mode, const int *two, const int *board, int nb) {
    int c[16];
    c[0] = two[0]; c[1] = two[1];
    for (int i = 0; i < nb; i++) c[2 + i] = board[i];
    if (mode == 1) return lit_category(c, nb + 2);
    return -treys_evaluate(c, nb + 2); /* negate: "greater is better" like the literal ranker */
}

/* ---- HandCalculations.py:157-176 HandStrength -> (ahead, tied, behind) ---------------------------- */
void ehs_hs_counts(int mode, const int *our, const int *board, int nb, uint32_t *out3) {
    int known[8], deck[52];
    known[0] = our[0]; known[1] = our[1];
    for (int i = 0; i < nb; i++) known[2 + i] = board[i];
    int nd = deck_without(known, nb + 2, deck); /* :148-150 */
    int ourrank = rank_of(mode, our, board, nb); /* :163 */
    uint32_t ahead = 0, tied = 0, behind = 0;
    for (int i = 0; i < nd; i++)
        for (int j = i + 1; j < nd; j++) { /* :141-145, :166 */
            int opp[2] = {deck[i], deck[j]};
            int opprank = rank_of(mode, opp, board, nb);
            if (ourrank > opprank) ahead++;
            else if (ourrank == opprank) tied++;
            else behind++;
        }
    out3[0] = ahead; out3[1] = tied; out3[2] = behind;
}

/* ---- HandCalculations.py:178-226 HandPotential -> HP[3][3] (row-major), HPTotal[3] ------------------
 * mode 1: literal.  Runouts = all pairs of deck \ board (:152-154, :207): they may contain our hole
 *         cards or the opponent's; always two cards, whatever the street.
 * mode 0: canonical.  Runouts = cards still needed to complete a 5-card board, drawn from the cards
 *         unseen by both players: flop -> C(45,2) pairs, turn -> 44 single rivers, river -> none. */
void ehs_hp_counts(int mode, const int *our, const int *board, int nb, uint32_t *hp9, uint32_t *hptotal3) {
    int known[8], deck[52];
    memset(hp9, 0, 9 * sizeof(uint32_t));
    memset(hptotal3, 0, 3 * sizeof(uint32_t));
    known[0] = our[0]; known[1] = our[1];
    for (int i = 0; i < nb; i++) known[2 + i] = board[i];
    int nd = deck_without(known, nb + 2, deck);
    int ourrank = rank_of(mode, our, board, nb); /* :190 */

    int rdeck[52], nr = 0;
    if (mode == 1) nr = deck_without(board, nb, rdeck); /* :153 */

    for (int i = 0; i < nd; i++)
        for (int j = i + 1; j < nd; j++) { /* :193 */
            int opp[2] = {deck[i], deck[j]};
            int opprank = rank_of(mode, opp, board, nb); /* :194 */
            int index = ourrank > opprank ? 0 : (ourrank == opprank ? 1 : 2); /* :196-201 */
            hptotal3[index]++;                                               /* :203 */
            int nboard[8];
            for (int k = 0; k < nb; k++) nboard[k] = board[k];
            if (mode == 1) {
                for (int a = 0; a < nr; a++)
                    for (int b = a + 1; b < nr; b++) { /* :207-209 */
                        nboard[nb] = rdeck[a]; nboard[nb + 1] = rdeck[b];
                        int ourbest = rank_of(1, our, nboard, nb + 2); /* :210 */
                        int oppbest = rank_of(1, opp, nboard, nb + 2); /* :211 */
                        int o = ourbest > oppbest ? 0 : (ourbest == oppbest ? 1 : 2);
                        hp9[index * 3 + o]++; /* :213-218 */
                    }
            } else if (nb == 3) {
                for (int a = 0; a < nd; a++) {
                    if (a == i || a == j) continue;
                    for (int b = a + 1; b < nd; b++) {
                        if (b == i || b == j) continue;
                        nboard[3] = deck[a]; nboard[4] = deck[b];
                        int ourbest = rank_of(0, our, nboard, 5);
                        int oppbest = rank_of(0, opp, nboard, 5);
                        int o = ourbest > oppbest ? 0 : (ourbest == oppbest ? 1 : 2);
                        hp9[index * 3 + o]++;
                    }
                }
            } else if (nb == 4) {
                for (int a = 0; a < nd; a++) {
                    if (a == i || a == j) continue;
                    nboard[4] = deck[a];
                    int ourbest = rank_of(0, our, nboard, 5);
                    int oppbest = rank_of(0, opp, nboard, 5);
                    int o = ourbest > oppbest ? 0 : (ourbest == oppbest ? 1 : 2);
                    hp9[index * 3 + o]++;
                }
            }
        }
}

/* Same counts as ehs_hp_counts(mode 0, flop/turn) but with our final rank cached per runout
 * (SURVEY.md App. A.3); checked equal to the loop above in tests/test_oracle_ehs.py and used where
 * the CPU suite needs many canonical situations quickly. */
void ehs_hp_counts_treys_cached(const int *our, const int *board, int nb, uint32_t *hp9, uint32_t *hptotal3) {
    int known[8], deck[52];
    static __thread int16_t ourcache[52][52];
    memset(hp9, 0, 9 * sizeof(uint32_t));
    memset(hptotal3, 0, 3 * sizeof(uint32_t));
    known[0] = our[0]; known[1] = our[1];
    for (int i = 0; i < nb; i++) known[2 + i] = board[i];
    int nd = deck_without(known, nb + 2, deck);
    int ourrank = rank_of(0, our, board, nb);
    int nboard[8];
    for (int k = 0; k < nb; k++) nboard[k] = board[k];
    if (nb == 3) {
        for (int a = 0; a < nd; a++)
            for (int b = a + 1; b < nd; b++) {
                nboard[3] = deck[a]; nboard[4] = deck[b];
                ourcache[a][b] = (int16_t)rank_of(0, our, nboard, 5);
            }
    } else if (nb == 4) {
        for (int a = 0; a < nd; a++) {
            nboard[4] = deck[a];
            ourcache[a][a] = (int16_t)rank_of(0, our, nboard, 5);
        }
    }
    for (int i = 0; i < nd; i++)
        for (int j = i + 1; j < nd; j++) {
            int opp[2] = {deck[i], deck[j]};
            int opprank = rank_of(0, opp, board, nb);
            int index = ourrank > opprank ? 0 : (ourrank == opprank ? 1 : 2);
            hptotal3[index]++;
            if (nb == 3) {
                for (int a = 0; a < nd; a++) {
                    if (a == i || a == j) continue;
                    for (int b = a + 1; b < nd; b++) {
                        if (b == i || b == j) continue;
                        nboard[3] = deck[a]; nboard[4] = deck[b];
                        int ourbest = ourcache[a][b];
                        int oppbest = rank_of(0, opp, nboard, 5);
                        hp9[index * 3 + (ourbest > oppbest ? 0 : (ourbest == oppbest ? 1 : 2))]++;
                    }
                }
            } else if (nb == 4) {
                for (int a = 0; a < nd; a++) {
                    if (a == i || a == j) continue;
                    nboard[4] = deck[a];
                    int ourbest = ourcache[a][a];
                    int oppbest = rank_of(0, opp, nboard, 5);
                    hp9[index * 3 + (ourbest > oppbest ? 0 : (ourbest == oppbest ? 1 : 2))]++;
                }
            }
        }
}

/* Flop counts of mode 0 once more, now with BOTH final ranks cached: ours per runout (as above) and the
 * opponent's per 4-set of unseen cards -- the opponent's final hand is board + holding + runout, which
 * depends only on the union of the two pairs, so the C(47,4) = 178,365 distinct seven-card hands are
 * evaluated once each instead of 1081 x 990 = 1,070,190 times.  The loop itself keeps the reference's
 * shape (HandCalculations.py:193-218: holdings outside, runouts inside).  Checked equal to
 * ehs_hp_counts(mode 0) in tests/test_oracle_ehs.py; this is what lets the GPU suite compare ALL
 * 65,536 bench situations with the oracle. */
static int choose_tab[48][5];
static void choose_init(void) {
    if (choose_tab[4][4]) return;
    for (int n = 0; n < 48; n++) {
        choose_tab[n][0] = 1;
        for (int k = 1; k < 5; k++) choose_tab[n][k] = n == 0 ? 0 : choose_tab[n - 1][k - 1] + choose_tab[n - 1][k];
    }
}
void ehs_hp_counts_treys_4set(const int *our, const int *board, uint32_t *hp9, uint32_t *hptotal3) {
    int known[8], deck[52];
    static __thread int16_t ourcache[47][47];
    static __thread int16_t oppcache[178365];
    static __thread uint8_t cls[47][47];
    choose_init();
    memset(hp9, 0, 9 * sizeof(uint32_t));
    memset(hptotal3, 0, 3 * sizeof(uint32_t));
    known[0] = our[0]; known[1] = our[1];
    for (int i = 0; i < 3; i++) known[2 + i] = board[i];
    int nd = deck_without(known, 5, deck);
    if (nd != 47) return; /* malformed situation: all-zero counts */
    int ourrank = rank_of(0, our, board, 3);
    int c7[7] = {board[0], board[1], board[2], 0, 0, 0, 0};
    for (int w = 0; w < nd; w++)
        for (int x = w + 1; x < nd; x++) {
            int two[2] = {deck[w], deck[x]};
            int nboard[5] = {board[0], board[1], board[2], deck[w], deck[x]};
            ourcache[w][x] = (int16_t)rank_of(0, our, nboard, 5);      /* :210 */
            int opprank = rank_of(0, two, board, 3);                   /* :194 */
            cls[w][x] = (uint8_t)(ourrank > opprank ? 0 : (ourrank == opprank ? 1 : 2));
            hptotal3[cls[w][x]]++;                                     /* :203 */
            for (int y = x + 1; y < nd; y++)
                for (int z = y + 1; z < nd; z++) {
                    c7[3] = deck[w]; c7[4] = deck[x]; c7[5] = deck[y]; c7[6] = deck[z];
                    oppcache[w + choose_tab[x][2] + choose_tab[y][3] + choose_tab[z][4]] = (int16_t)-treys_evaluate(c7, 7);
                }
        }
    for (int i = 0; i < nd; i++)
        for (int j = i + 1; j < nd; j++) {                             /* :193 */
            int index = cls[i][j];
            for (int a = 0; a < nd; a++) {
                if (a == i || a == j) continue;
                for (int b = a + 1; b < nd; b++) {                     /* :207-209 */
                    if (b == i || b == j) continue;
                    /* sorted union of {i<j} and {a<b} */
                    int w = i < a ? i : a, z = j > b ? j : b;
                    int m1 = i < a ? a : i, m2 = j > b ? b : j;
                    int x = m1 < m2 ? m1 : m2, y = m1 < m2 ? m2 : m1;
                    int ourbest = ourcache[a][b];
                    int oppbest = oppcache[w + choose_tab[x][2] + choose_tab[y][3] + choose_tab[z][4]];
                    hp9[index * 3 + (ourbest > oppbest ? 0 : (ourbest == oppbest ? 1 : 2))]++; /* :213-218 */
                }
            }
        }
}

/* ---- batch drivers over packed situations uint8 [N,8]: h0 h1 b0..b4 (0xFF = absent) pad ------------
 * out rows are uint32 [16]: ahead tied behind | HP[9] row-major | HPTotal[3] | n_board  (include/holdem_b200.h) */
static int unpack(const uint8_t *row, int *our, int *board) {
    our[0] = row[0]; our[1] = row[1];
    int nb = 0;
    for (int k = 2; k < 7; k++)
        if (row[k] != 0xFF) board[nb++] = row[k];
    return nb;
}

void ehs_hs_batch(int mode, const uint8_t *sit, int64_t n, uint32_t *out /* [n,3] */) {
    treys_init();
    for (int64_t s = 0; s < n; s++) {
        int our[2], board[5];
        int nb = unpack(sit + 8 * s, our, board);
        ehs_hs_counts(mode, our, board, nb, out + 3 * s);
    }
}

/* cache: 0 = our rank cached per runout only (one opponent evaluation per (holding, runout) pair, the
 * reference's loop shape -- what the CPU baseline times), 1 = also the opponent's rank per 4-set (flop rows). */
void ehs_hp_batch_ex(int mode, int cache, const uint8_t *sit, int64_t n, uint32_t *out /* [n,16] */) {
    treys_init();
    for (int64_t s = 0; s < n; s++) {
        int our[2], board[5];
        int nb = unpack(sit + 8 * s, our, board);
        uint32_t *o = out + 16 * s;
        if (mode == 0 && nb == 3 && cache) ehs_hp_counts_treys_4set(our, board, o + 3, o + 12);
        else if (mode == 0) ehs_hp_counts_treys_cached(our, board, nb, o + 3, o + 12);
        else ehs_hp_counts(1, our, board, nb, o + 3, o + 12);
        o[0] = o[12]; o[1] = o[13]; o[2] = o[14]; /* HPTotal == HS counts (same loop, :168-173 vs :196-203) */
        o[15] = (uint32_t)nb;
    }
}

void ehs_hp_batch(int mode, const uint8_t *sit, int64_t n, uint32_t *out /* [n,16] */) {
    ehs_hp_batch_ex(mode, 0, sit, n, out);
}

void lit_category_batch(const uint8_t *rows, int ncards, int stride, int64_t n, uint8_t *out) {
    for (int64_t s = 0; s < n; s++) {
        int c[16];
        for (int k = 0; k < ncards; k++) c[k] = rows[s * stride + k];
        out[s] = (uint8_t)lit_category(c, ncards);
    }
}
