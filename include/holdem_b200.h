/*
 * holdem_b200.h -- C ABI of the B200-native hand-strength / hand-potential path.
 *
 * Drop-in boundary for dsowsy/dask-holdem-simulator's HandCalculations.py (reference file:line cited per
 * entry point).  Plain pointers and sizes only: no torch, no C++ types.  The shared library is
 * dask-holdem-simulator_b200/holdem_b200/libholdem_b200.so; the Python shim binds it with ctypes
 * (INTEGRATION.md shows the stub).  There is NO CPU fallback: every compute entry point needs a CUDA
 * device of compute capability 10.x and returns HOLDEM_ERR_NO_DEVICE otherwise.
 *
 * Card encoding (HandCalculations.py:53-54): card in [0,52), rank = card % 13 (0 = deuce .. 12 = ace),
 * suit = card / 13 (0 c, 1 d, 2 h, 3 s).
 *
 * Situation row, uint8[8]:  h0 h1 b0 b1 b2 b3 b4 pad   -- absent board cards and pad are 0xFF; the board
 * holds 3, 4 or 5 cards, contiguous from b0; all cards of a row are distinct.
 *
 * Result row of the HP entry points, uint32[16] (HOLDEM_HP_ROW):
 *   [0..2]   ahead, tied, behind            HandStrength counts     (HandCalculations.py:168-173)
 *   [3..11]  HP[3][3] row-major             HP[index][outcome]      (HandCalculations.py:213-218)
 *   [12..14] HPTotal[3]                     per-index opponent count (HandCalculations.py:203)
 *   [15]     number of board cards; 0xFFFFFFFF in every column marks a malformed input row
 *
 * mode: HOLDEM_MODE_TREYS      canonical Effective-Hand-Strength enumeration on Treys ranks (1..7462, lower
 *                              is better): opponents C(47,2)/C(46,2)/C(45,2); runouts are the cards still to
 *                              come drawn from the cards unseen by both players (flop 990 pairs per opponent,
 *                              turn 44 rivers, river none).
 *       HOLDEM_MODE_REFERENCE  HandCalculations.py exactly as written: 9-category ranker
 *                              CalculateHandRankValue (:45-82), runouts = ALL pairs of deck \ board
 *                              (:152-154; 1176/1128/1081), multiset semantics and quirks included.
 *
 * Every *_batch entry point is stream-ordered on `stream` (a cudaStream_t passed as void*; NULL = the legacy
 * default stream), takes DEVICE pointers and does not synchronise.  Every *_host entry point takes HOST
 * pointers, performs the host<->device copies itself and returns after the results are in host memory.
 * Return value: 0 on success, a negative HOLDEM_ERR_* otherwise; holdem_last_error() gives the message of the
 * calling thread's last failure.  Nothing throws across this boundary.
 */
#ifndef HOLDEM_B200_H
#define HOLDEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOLDEM_ABI_VERSION 2

#define HOLDEM_MODE_TREYS 0
#define HOLDEM_MODE_REFERENCE 1

#define HOLDEM_HP_ROW 16
#define HOLDEM_HS_ROW 3
#define HOLDEM_BAD_ROW 0xFFFFFFFFu

#define HOLDEM_OK 0
#define HOLDEM_ERR_NO_DEVICE (-1)   /* no CUDA device / not sm_100 / driver error at init            */
#define HOLDEM_ERR_BAD_ARG (-2)     /* null pointer, negative size, unknown mode, unsupported n_cards */
#define HOLDEM_ERR_CUDA (-3)        /* a CUDA runtime call failed; see holdem_last_error()            */
#define HOLDEM_ERR_NOT_INIT (-4)    /* holdem_init() has not succeeded on the current device          */
#define HOLDEM_ERR_BAD_INPUT (-5)   /* *_host only: at least one malformed row (marked HOLDEM_BAD_ROW) */
#define HOLDEM_ERR_INTERNAL (-6)    /* table self-check failed                                        */

/* ---- life cycle ------------------------------------------------------------------------------------ */

int holdem_abi_version(void);
const char *holdem_last_error(void);

/* Builds the lookup tables on the host (once per process) and uploads them to `device` (once per device).
 * Idempotent and thread-safe.  Replaces the module-level `evaluator = Evaluator()` of
 * HandCalculations.py:268 and the per-call `Evaluator()` of :231 (Treys builds its tables there). */
int holdem_init(int device);

/* Host-only consistency check of the generated tables (7462 classes, perfect hashes collision-free,
 * category sizes).  Needs no GPU.  Returns HOLDEM_OK or HOLDEM_ERR_INTERNAL. */
int holdem_selftest_tables(void);

/* Number of streaming multiprocessors of the initialised device (grid sizing for callers), or <0. */
int holdem_sm_count(int device);

/* ---- batched evaluator: Treys rank ------------------------------------------------------------------
 * Replaces treys.Evaluator.evaluate as reached from evaluate_best_hand (HandCalculations.py:229-240) and
 * CalculateHandValue (pokerlib_to_treys.py:15-21).  cards: uint8[n,8], the first n_cards (5, 6 or 7)
 * columns of each row are the hand, the rest is ignored.  rank_out: uint16[n], 1..7462 (0 = malformed row).
 * Malformed = a card > 51, always detected.  Duplicate cards inside a row are detected (rank 0) by
 * holdem_eval_batch_strict and by the _host entry; the plain device entry, like Treys itself, does not look for them
 * in its large-batch fast path and returns an unspecified rank for such rows. */
int holdem_eval_batch(const uint8_t *d_cards, int n_cards, int64_t n, uint16_t *d_rank_out, void *stream);
int holdem_eval_batch_strict(const uint8_t *d_cards, int n_cards, int64_t n, uint16_t *d_rank_out, void *stream);
int holdem_eval_batch_host(const uint8_t *h_cards, int n_cards, int64_t n, uint16_t *h_rank_out);

/* Exhaustive validation sweep (BASELINE.json configs[1]): enumerates all C(52,7) = 133,784,560 seven-card
 * hands in-kernel.  d_hist: uint64[7463] (index = rank), d_sums: uint64[2] = {sum rank, sum rank^2}.
 * Both are overwritten. */
int holdem_exhaustive7(uint64_t *d_hist, uint64_t *d_sums, void *stream);
int holdem_exhaustive7_host(uint64_t *h_hist, uint64_t *h_sums);

/* ---- batched literal ranker ---------------------------------------------------------------------------
 * Replaces CalculateHandRankValue (HandCalculations.py:45-82) on the concatenation ourcards + boardcards.
 * cards: uint8[n,row_stride], first n_cards (2..9) columns used, duplicates allowed (multiset semantics).
 * cat_out: uint8[n], 0..8, higher is better (0xFF = malformed row: a card > 51). */
int holdem_category_batch(const uint8_t *d_cards, int n_cards, int row_stride, int64_t n, uint8_t *d_cat_out,
                          void *stream);
int holdem_category_batch_host(const uint8_t *h_cards, int n_cards, int row_stride, int64_t n,
                               uint8_t *h_cat_out);

/* ---- hand strength -------------------------------------------------------------------------------------
 * Replaces the loop of HandStrength (HandCalculations.py:157-176).  sit: uint8[n,8] situation rows;
 * out: uint32[n,3] = ahead, tied, behind.  The float is derived by the caller exactly as :175. */
int holdem_hs_batch(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *stream);
int holdem_hs_batch_host(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out);

/* ---- hand strength + hand potential ----------------------------------------------------------------------
 * Replaces the loops of HandPotential (HandCalculations.py:178-226) and, as a by-product, HandStrength.
 * sit: uint8[n,8]; out: uint32[n,HOLDEM_HP_ROW].  PPot/NPot are derived by the caller exactly as :221-224. */
int holdem_hp_batch(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *stream);
int holdem_hp_batch_host(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out);
/* The same through the chunked copy pipeline (16,384-row chunks, library-side pinned staging, two streams) even when the
 * caller's buffers are page-locked; holdem_hp_batch_host uses it for pageable buffers. */
int holdem_hp_batch_host_staged(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out);

/* holdem_hp_batch that also stores every result row into up to 8 other GPUs' buffers (peer memory mapped into this process:
 * CUDA IPC / symmetric memory; NVLink stores issued by the kernels themselves, no separate collective).  peer_out is a HOST
 * array of n_peers device pointers, peer_out[k] pointing at the row of peer k's [N,16] buffer that corresponds to row 0 of
 * this call (i.e. already offset by this rank's first row).  The caller orders the peers (a barrier before the buffers are
 * overwritten and one after this call's stream work); HandCalculations' multi-GPU layer (holdem_b200/sharded.py) does. */
int holdem_hp_batch_peers(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *const *peer_out, int n_peers,
                          void *stream);

/* Whole tables: every seat of a hand shares the board (GameBoard.py:12 MAX_PLAYERS = 22; GameBoard.py:204-229 walks the
 * seats one by one).  d_hole uint8[hands,seats,2], d_board uint8[hands,5] (the first n_board = 3, 4 or 5 cards are used),
 * d_out uint32[hands,seats,HOLDEM_HP_ROW]: the rows holdem_hp_batch would return for situation (hole[h][s], board[h][:n_board]).
 * The seats of one board are adjacent in the work order, so the per-board table rows are fetched once per board. */
int holdem_table_hp_batch(const uint8_t *d_hole, const uint8_t *d_board, int64_t hands, int seats, int n_board, int mode,
                          uint32_t *d_out, void *stream);

/* Showdown of whole tables in one pass (GameBoard.determine_winner, GameBoard.py:204-229, which evaluates seat by seat through
 * evaluate_best_hand): d_hole uint8[hands,seats,2], d_board uint8[hands,5], d_active optional uint8[hands,seats] (NULL = every
 * seat is in the hand).  d_ranks uint16[hands,seats]: Treys rank of every seat (0 = malformed seat: bad or duplicate card);
 * d_winners uint32[hands,2]: bit s of word 0 = seat s holds the BEST hand among the active seats (lowest rank: the correct
 * winner, ties share), of word 1 = seat s is what the reference picks (GameBoard.py:210 takes max() of the Treys values, i.e.
 * the worst hand). */
int holdem_table_showdown(const uint8_t *d_hole, const uint8_t *d_board, const uint8_t *d_active, int64_t hands, int seats,
                          uint16_t *d_ranks, uint32_t *d_winners, void *stream);

/* Same results as holdem_hp_batch computed by the plain (non-deduplicating) enumeration kernel: one opponent
 * evaluation per (opponent holding, runout) pair.  Kept as the independent second implementation the parity
 * tests cross-check the deduplicating flop / turn kernels against; it is also what holdem_hp_batch itself runs for
 * HOLDEM_MODE_REFERENCE and for Treys-mode river rows. */
int holdem_hp_batch_direct(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *stream);

/* Treys-mode flop rows through one named implementation of the flop kernel (A/B measurements and cross-checks; every
 * variant returns the same counts): 4 = rank-multiset kernel (what holdem_hp_batch runs), 5 = the same kernel without the
 * closed-form sum of the flush steps' add parts, 6 = the production kernel working the rows in input order instead of heaviest
 * board texture first, 7 = the same kernel with the undo half of the flush part walked step by step instead of taken as weighted
 * population counts over the generic part's compare bits (the construction of round 1), 3 = deduplicating 4-set walk,
 * 2 = the earlier warp-per-pair 4-set kernel.  Rows with 4- or 5-card boards are left untouched. */
int holdem_hp_flop_variant(const uint8_t *d_sit, int64_t n, int variant, uint32_t *d_out, void *stream);

/* The same for Treys-mode turn rows (four board cards): 4 = rank-multiset kernel (production), 6 = the same working the rows in
 * input order with static striding instead of heaviest board texture first, 3 = deduplicating 3-set walk.
 * Rows of other streets are left untouched. */
int holdem_hp_turn_variant(const uint8_t *d_sit, int64_t n, int variant, uint32_t *d_out, void *stream);

/* Whole Treys-mode batches (all streets) through named implementations: flop 2..7 and turn 4 / 3 as above, river 0 =
 * rank-pair HS kernel in its HP-row form (production), 1 = plain enumeration.  holdem_hp_batch == (4, 4, 0).  These entry
 * points replace the HOLDEM_*_KERNEL environment switches of ABI version 1: production dispatch reads no environment. */
int holdem_hp_batch_variant(const uint8_t *d_sit, int64_t n, int flop_variant, int turn_variant, int river_variant,
                            uint32_t *d_out, void *stream);
/* HOLDEM_MODE_REFERENCE batches (all streets) through a named kernel: 5 = rank-multiset kernel (what holdem_hp_batch runs),
 * 4 = the earlier 4-set walk with arithmetic ranking of the irregular runouts.  Same counts. */
int holdem_hp_ref_variant(const uint8_t *d_sit, int64_t n, int variant, uint32_t *d_out, void *stream);
/* holdem_hs_batch through a named kernel: 0 = production (Treys mode: rank-pair kernel), 1 = one evaluation per holding. */
int holdem_hs_batch_variant(const uint8_t *d_sit, int64_t n, int mode, int variant, uint32_t *d_out, void *stream);
/* holdem_exhaustive7 through a named kernel: 0 = tuple kernel (production), 1 = chunk-per-thread kernel. */
int holdem_exhaustive7_variant(int variant, uint64_t *d_hist, uint64_t *d_sums, void *stream);

/* The production flop kernel with its phase clocks switched on: d_cycles uint64[8] receives the SM cycles one thread of each
 * situation group spent in [0] card lists + small table rows, [1] hand strength + rank-pair sets, [4] the entries of the runouts
 * that complete our flush, [5] step descriptors, [2] next claim + wait for the T row, [6] generic tasks up to the group barrier,
 * [3] flush tasks + reduction ([7] unused), summed over all situations.  Profiling aid; d_out receives the normal rows. */
int holdem_hp_flop_phase_cycles(const uint8_t *d_sit, int64_t n, uint32_t *d_out, uint64_t *d_cycles, void *stream);

/* ---- callers and data formats either side of the path -------------------------------------------
 * Position-dependent 64-bit content hash of a [n, words_per_row] uint32 buffer: sum over rows r of
 * mix(fnv1a64(row r, seed first_row + r)).  Equal hashes across GPU counts / kernel variants mean equal rows (a plain sum
 * of the counts is a structural constant).  d_hash: uint64[1], overwritten.  Shards combine by addition. */
int holdem_rows_hash(const uint32_t *d_rows, int64_t n, int words_per_row, int64_t first_row, uint64_t *d_hash, void *stream);

/* Batched Decision.decision_based_on_blind_position (Decision.py:4-48).  Positions: 0 'SB', 1 'BB', 2 'NONE' -- one per row
 * (d_pos uint8[n]) or pos_all for the whole batch when d_pos is NULL (anything else: HOLDEM_ERR_BAD_ARG with the reference's
 * message, Decision.py:8-9; a bad per-row code gives action -2).  d_action int8[n]: 0 fold, 1 check, 2 call, 3 raise,
 * -1 where the reference falls through and returns None (hand_strength > 0.8 and npot >= 0.3). */
int holdem_decide_batch(const double *d_hs, const double *d_ppot, const double *d_npot, const uint8_t *d_pos, int pos_all,
                        int64_t n, int8_t *d_action, void *stream);
/* The same straight from HS/HP result rows (uint32[n,16]): HS as HandCalculations.py:175 (double), PPot/NPot as :221-224 with
 * each HP row divided by the runouts per opponent (probabilities; 0 where undefined), then the tree above -- one fused pass
 * over the rows.  d_hs_ppot_npot: optional double[n,3] receiving the three inputs of the tree.  Malformed rows give -2. */
int holdem_decide_from_counts(const uint32_t *d_rows, const uint8_t *d_pos, int pos_all, int64_t n, int8_t *d_action,
                              double *d_hs_ppot_npot, void *stream);

/* utils.card_to_int wire format (utils.py:82-85: rank * 4 + suit, what GameBoard / Player publish) -> situation rows in the
 * path encoding (HandCalculations.py:53-54).  d_hole int64[n,2], d_board int64[n,n_board], n_board 3..5; d_sit uint8[n,8].
 * A value outside [0,52) is written as 0xFE, which every kernel reports as a malformed row. */
int holdem_pack_rank_major(const int64_t *d_hole, const int64_t *d_board, int n_board, int64_t n, uint8_t *d_sit, void *stream);

/* Kernels launched by this library in this process so far (every launch site counts itself). */
unsigned long long holdem_kernel_launches(void);

/* ---- roofline micro-benchmarks (bench.py) ---------------------------------------------------------------
 * Issues `iters` dependent uniformly-random uint16 gathers per thread over a shared-memory table of
 * `table_bytes` from every lane of a full grid and returns the elapsed device time (CUDA events on `stream`).
 * pattern 0: random indices (bank conflicts as they fall); 1: conflict-free (lane-strided) indices. */
int holdem_microbench_smem_gather(int table_bytes, int iters, int pattern, int blocks_per_sm, int threads,
                                  float *elapsed_ms, double *lookups, void *stream);

/* Integer issue-rate micro-benchmark: every thread runs eight independent chains of 16 instructions per iteration.
 * mix 0: compare (ALU pipe) + predicated multiply-add (FMA pipe) pairs, the instruction pair of the HP kernels;
 * mix 1: multiply-adds only; mix 2: logic ops only.  Returns elapsed device time and the warp instructions issued. */
int holdem_microbench_int_issue(int iters, int mix, int blocks_per_sm, int threads, float *elapsed_ms,
                                double *warp_instructions, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HOLDEM_B200_H */
