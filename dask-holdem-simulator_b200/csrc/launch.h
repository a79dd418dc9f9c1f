// launch.h -- host-callable launchers of the kernels (internal to the library; the public surface is
// include/holdem_b200.h).
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

namespace holdem {

struct DevTables;

// Result rows can be stored into other GPUs' buffers as well (peer memory over NVLink): p[k] points at the row that
// corresponds to the kernel's row 0 inside peer k's [N,16] buffer.  n = 0: local output only.
struct PeerOut {
    uint32_t *p[8];
    int n;
};

constexpr uint32_t kFlushEntries = 8192;

// Every kernel launch of the library is counted (process-wide); holdem_kernel_launches() reports the total, so callers
// (bench.py's gpu_launches) report a measured count instead of a formula.
extern std::atomic<unsigned long long> g_kernel_launches;
inline void count_launch() { g_kernel_launches.fetch_add(1ull, std::memory_order_relaxed); }

// strict: also detect duplicate cards inside a row (builds the card set; ~1.5x the instructions of the fast path).
cudaError_t launch_eval_batch(const DevTables &T, const uint8_t *cards, int n_cards, int64_t n, uint16_t *out,
                              int sms, bool strict, cudaStream_t st);
// Whole-table showdown: ranks[hands, seats] and per hand two seat bit masks (best = lowest rank, reference rule = highest rank).
cudaError_t launch_table_showdown(const DevTables &T, const uint8_t *hole, const uint8_t *board, const uint8_t *active,
                                  int64_t hands, int seats, uint16_t *ranks, uint32_t *winners, int sms, cudaStream_t st);
cudaError_t launch_category_batch(const uint8_t *cards, int n_cards, int row_stride, int64_t n, uint8_t *out,
                                  int sms, cudaStream_t st);
// variant 0 = tuple kernel (production), 1 = the older chunk-per-thread kernel.
cudaError_t launch_exhaustive7(const DevTables &T, uint64_t *hist, uint64_t *sums, int sms, cudaStream_t st, int variant = 0);
cudaError_t launch_smem_gather(int table_bytes, int iters, int pattern, int blocks, int threads, cudaStream_t st);
// 16 warp instructions per thread per iteration; mix 0 = compare + predicated multiply-add, 1 = FMA pipe only, 2 = ALU pipe only
cudaError_t launch_int_issue(int iters, int mix, int blocks, int threads, cudaStream_t st);

// variant 0 = production (rank-pair kernels of both modes), 1 = the per-holding kernel for both modes.
cudaError_t launch_hs_batch(const DevTables &T, const uint8_t *sit, int64_t n, int mode, uint32_t *out, int sms,
                            cudaStream_t st, int variant = 0);
// Treys-mode hand strength from rank pairs (kernels_hs4.cu): same results, what launch_hs_batch runs for HOLDEM_MODE_TREYS
// unless HOLDEM_HS_KERNEL=1 selects the holding-by-holding kernel.
cudaError_t launch_hs_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, cudaStream_t st);
// The same kernel as the last pass of holdem_hp_batch (treys mode): HP rows (hand strength only) for the five-card-board rows.
cudaError_t launch_hp_river_rows(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, int *run_flag,
                                 cudaStream_t st);
// run_flag: optional device int; the kernel exits at once unless bit 1 is set.  skip_boards: leave 3- and 4-card-board rows
// alone (they belong to the flop / turn kernels).
cudaError_t launch_hp_direct(const DevTables &T, const uint8_t *sit, int64_t n, int mode, uint32_t *out, int sms,
                             int *run_flag, int skip_boards, cudaStream_t st);
// Treys-mode turn rows only; exits at once unless bit 0 of *run_flag is set (the flop kernel saw 4-card boards).
cudaError_t launch_hp_turn_treys(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                 const int *run_flag, cudaStream_t st);
// Treys-mode flop rows only; sets bit 0 / bit 1 of *other_rows_flag when it met turn / river rows (left to the kernels below).
// impl: 3 = lane-per-(a,b,d) kernel (default), 2 = the earlier warp-per-(a,b) kernel (kept for A/B measurements).
cudaError_t launch_hp_flop_treys(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                 int *other_rows_flag, int impl, cudaStream_t st);
// Rank-multiset flop kernel (kernels_hp_flop4.cu): same contract and bit-identical counts, the default.
// closed_form_add: sum the add parts of the flush steps in closed form where no runout of the warp gives us a flush or better
// (production); false = walk them step by step (holdem_hp_flop_variant 5, kept for A/B and as a cross-check).
// bit_parallel_undo: the undo half of the flush part as weighted population counts over the compare bits the generic tasks leave
// behind (production); false = the step-by-step undo half of the earlier kernel (holdem_hp_flop_variant 7).
// prof: optional device uint64[4], cycles spent per phase summed over the situation groups (tables / W+T / descriptors /
// tasks); nullptr in production.
cudaError_t launch_hp_flop_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                    int *other_rows_flag, unsigned long long *prof, const PeerOut &peers, cudaStream_t st,
                                    bool closed_form_add = true, uint32_t *order_scratch = nullptr, bool bit_parallel_undo = true);
// order_scratch: optional device buffer filled by launch_street_order (kernels_order.cu): the situations are then worked heaviest
// board texture first (results still land by row number).
constexpr uint32_t kOrderHeader = 128;        // words of counters in front of the two row lists
size_t street_order_scratch_bytes(int64_t n);
cudaError_t launch_street_order(const uint8_t *sit, int64_t n, uint32_t *scratch, int sms, cudaStream_t st);

// Rank-multiset turn kernel (kernels_hp_turn4.cu): same contract as launch_hp_turn_treys, bit-identical counts, the default.
cudaError_t launch_hp_turn_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                    int *run_flag, const PeerOut &peers, cudaStream_t st, const uint32_t *order_scratch = nullptr);
cudaError_t launch_hp_turn_treys_v4_counter(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                            int *slot, const PeerOut &peers, cudaStream_t st, const uint32_t *order_scratch);
// Copies result rows of `out` to the peers: all rows (run_flag == nullptr) or, when run_flag is given, the five-card-board rows
// and only if bit 1 of *run_flag is set (the rows the plain-enumeration kernel produced).
cudaError_t launch_rows_to_peers(const uint8_t *sit, int64_t n, const uint32_t *out, const PeerOut &peers, int sms,
                                 const int *run_flag, cudaStream_t st);

// HOLDEM_MODE_REFERENCE, all streets: 4-set walk for the regular runouts + arithmetic ranking of the irregular ones.
// bad_flag: optional device int, bit 2 is set when a malformed row was marked (every HP kernel does this on its flag word).
cudaError_t launch_hp_ref(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, int *bad_flag,
                          cudaStream_t st);

// HOLDEM_MODE_REFERENCE, all streets, rank-multiset formulation (kernels_hp_ref5.cu): the default.  flag_slot = the call's 16-byte
// slot ([0] flag word, bit 2 = malformed row marked; [1] work counter, zeroed here); rows go to the peers as well.
cudaError_t launch_hp_ref5(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, int *flag_slot,
                           const PeerOut &peers, cudaStream_t st);

// Reference-mode hand strength from rank pairs (kernels_hp_ref5.cu): what launch_hs_batch runs for HOLDEM_MODE_REFERENCE.
cudaError_t launch_hs_ref5(const uint8_t *sit, int64_t n, uint32_t *out, int sms, cudaStream_t st);

// Device-side construction of the reference-mode category tables (k = 3, 4, 5 board cards); called once per device by holdem_init.
size_t ref_t_table_bytes(int k);
cudaError_t launch_ref_t_table(uint8_t *table, int k, int sms, cudaStream_t st);

// kernels_aux.cu: content hash of result rows, the Decision.py threshold tree, rank-major wire ints -> situation rows.
cudaError_t launch_rows_hash(const uint32_t *rows, int64_t n, int words, int64_t first_row, uint64_t *out, int sms,
                             cudaStream_t st);
cudaError_t launch_decide(const double *hs, const double *ppot, const double *npot, const uint8_t *pos, int pos_all,
                          int64_t n, int8_t *out, int sms, cudaStream_t st);
cudaError_t launch_decide_from_counts(const uint32_t *rows, const uint8_t *pos, int pos_all, int64_t n, int8_t *out,
                                      double *hs_ppot_npot, int sms, cudaStream_t st);
cudaError_t launch_table_rows(const uint8_t *hole, const uint8_t *board, int64_t hands, int seats, int n_board,
                              uint8_t *sit, int sms, cudaStream_t st);
cudaError_t launch_pack_rank_major(const int64_t *hole, const int64_t *board, int n_board, int64_t n, uint8_t *sit,
                                   int sms, cudaStream_t st);

}  // namespace holdem
