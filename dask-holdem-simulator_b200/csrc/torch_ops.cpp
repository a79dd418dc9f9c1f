// torch_ops.cpp -- thin TORCH_LIBRARY layer over the C ABI (include/holdem_b200.h): torch.ops.holdem.*
//
// SURVEY.md section 8b / BASELINE.json north_star: the Python entry points "call through a thin PyTorch C++/CUDA extension
// (C-ABI) into hand-written sm_100a kernels".  Each op checks its tensors, switches to the tensor's device
// (c10::cuda::CUDAGuard), takes torch's CURRENT stream of that device, allocates its outputs with torch and calls exactly one
// C-ABI entry point -- no kernel lives here and nothing synchronises.  There is no CPU implementation registered: calling an op
// with CPU tensors raises ("no CPU fallback").  Built into dask-holdem-simulator_b200/holdem_b200/libholdem_b200_torch.so and
// loaded with torch.ops.load_library (holdem_b200/_ops.py); the ctypes binding (holdem_b200/_lib.py) stays for callers
// without torch (Option B).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "../../include/holdem_b200.h"

namespace {

void check(int rc) { TORCH_CHECK(rc == HOLDEM_OK, "libholdem_b200 error ", rc, ": ", holdem_last_error()); }

// Device guard + library initialisation on the tensor's device; returns torch's current stream there.
struct Enter {
    c10::cuda::CUDAGuard guard;
    void *stream;
    explicit Enter(const at::Tensor &t) : guard(t.device()) {
        check(holdem_init(t.get_device()));
        stream = c10::cuda::getCurrentCUDAStream(t.get_device()).stream();
    }
};

void require(const at::Tensor &t, const char *name, at::ScalarType dtype, int64_t cols) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (this path has no CPU fallback)");
    TORCH_CHECK(t.scalar_type() == dtype && t.is_contiguous(), name, " must be a contiguous ", dtype, " tensor");
    TORCH_CHECK(cols < 0 || (t.dim() == 2 && t.size(1) == cols), name, " must have shape [N,", cols, "]");
}

// Replaces treys.Evaluator.evaluate (HandCalculations.py:229-240): uint8 [N,8] -> int16 [N] Treys rank 1..7462.
at::Tensor eval(const at::Tensor &cards, int64_t n_cards, bool strict) {
    require(cards, "cards", at::kByte, 8);
    Enter e(cards);
    at::Tensor out = at::empty({cards.size(0)}, cards.options().dtype(at::kShort));
    auto fn = strict ? holdem_eval_batch_strict : holdem_eval_batch;
    check(fn(cards.data_ptr<uint8_t>(), (int)n_cards, cards.size(0), (uint16_t *)out.data_ptr(), e.stream));
    return out;
}

// Replaces CalculateHandRankValue (HandCalculations.py:45-82): uint8 [N,S] -> uint8 [N] category 0..8.
at::Tensor category(const at::Tensor &cards, int64_t n_cards) {
    require(cards, "cards", at::kByte, -1);
    TORCH_CHECK(cards.dim() == 2 && cards.size(1) >= n_cards, "cards must be [N, >= n_cards]");
    Enter e(cards);
    at::Tensor out = at::empty({cards.size(0)}, cards.options());
    check(holdem_category_batch(cards.data_ptr<uint8_t>(), (int)n_cards, (int)cards.size(1), cards.size(0),
                                out.data_ptr<uint8_t>(), e.stream));
    return out;
}

// Replaces the loop of HandStrength (HandCalculations.py:157-176): uint8 [N,8] -> int32 [N,3] ahead, tied, behind.
at::Tensor hs(const at::Tensor &sit, int64_t mode) {
    require(sit, "sit", at::kByte, 8);
    Enter e(sit);
    at::Tensor out = at::empty({sit.size(0), HOLDEM_HS_ROW}, sit.options().dtype(at::kInt));
    check(holdem_hs_batch(sit.data_ptr<uint8_t>(), sit.size(0), (int)mode, (uint32_t *)out.data_ptr(), e.stream));
    return out;
}

void require_out(const at::Tensor &sit, const at::Tensor &out) {
    require(out, "out", at::kInt, HOLDEM_HP_ROW);
    TORCH_CHECK(out.size(0) == sit.size(0) && out.device() == sit.device(), "out must be [N,16] on the same device as sit");
}

// Replaces the loops of HandPotential (HandCalculations.py:178-226): uint8 [N,8] -> int32 [N,16] result rows.
void hp_out(const at::Tensor &sit, int64_t mode, at::Tensor out) {
    require(sit, "sit", at::kByte, 8);
    require_out(sit, out);
    Enter e(sit);
    check(holdem_hp_batch(sit.data_ptr<uint8_t>(), sit.size(0), (int)mode, (uint32_t *)out.data_ptr(), e.stream));
}

at::Tensor hp(const at::Tensor &sit, int64_t mode) {
    require(sit, "sit", at::kByte, 8);
    at::Tensor out = at::empty({sit.size(0), HOLDEM_HP_ROW}, sit.options().dtype(at::kInt));
    hp_out(sit, mode, out);
    return out;
}

// holdem_hp_batch_peers: the kernels also store every row into the peers' buffers (addresses of peer-mapped memory, e.g. the
// buffer_ptrs of a torch symmetric-memory handle, already offset to this rank's first row).
void hp_peers(const at::Tensor &sit, int64_t mode, at::Tensor out, at::IntArrayRef peer_ptrs) {
    require(sit, "sit", at::kByte, 8);
    require_out(sit, out);
    TORCH_CHECK(peer_ptrs.size() <= 8, "at most 8 peer buffers");
    Enter e(sit);
    std::vector<void *> ptrs;
    for (int64_t p : peer_ptrs) ptrs.push_back(reinterpret_cast<void *>(static_cast<uintptr_t>(p)));
    check(holdem_hp_batch_peers(sit.data_ptr<uint8_t>(), sit.size(0), (int)mode, (uint32_t *)out.data_ptr(),
                                ptrs.empty() ? nullptr : ptrs.data(), (int)ptrs.size(), e.stream));
}

// Board-shared table evaluation (GameBoard.py:204-229, MAX_PLAYERS = 22 at :12): hole uint8 [H,S,2], board uint8 [H,5]
// -> int32 [H,S,16] HS+HP rows for the first n_board cards of each board.
at::Tensor table_hp(const at::Tensor &hole, const at::Tensor &board, int64_t n_board, int64_t mode) {
    TORCH_CHECK(hole.is_cuda() && board.is_cuda(), "hole / board must be CUDA tensors (no CPU fallback)");
    TORCH_CHECK(hole.scalar_type() == at::kByte && hole.is_contiguous() && hole.dim() == 3 && hole.size(2) == 2,
                "hole must be a contiguous uint8 [H,S,2] tensor");
    TORCH_CHECK(board.scalar_type() == at::kByte && board.is_contiguous() && board.dim() == 2 && board.size(1) == 5 &&
                    board.size(0) == hole.size(0) && board.device() == hole.device(),
                "board must be a contiguous uint8 [H,5] tensor on the device of hole");
    Enter e(hole);
    at::Tensor out = at::empty({hole.size(0), hole.size(1), HOLDEM_HP_ROW}, hole.options().dtype(at::kInt));
    check(holdem_table_hp_batch(hole.data_ptr<uint8_t>(), board.data_ptr<uint8_t>(), hole.size(0), (int)hole.size(1),
                                (int)n_board, (int)mode, (uint32_t *)out.data_ptr(), e.stream));
    return out;
}

// Whole-table showdown (GameBoard.py:204-229) -> (ranks int16 [H,S], winners int32 [H,2]: seat bit masks, best / reference rule).
std::tuple<at::Tensor, at::Tensor> table_showdown(const at::Tensor &hole, const at::Tensor &board,
                                                  const c10::optional<at::Tensor> &active) {
    TORCH_CHECK(hole.is_cuda() && board.is_cuda(), "hole / board must be CUDA tensors (no CPU fallback)");
    TORCH_CHECK(hole.scalar_type() == at::kByte && hole.is_contiguous() && hole.dim() == 3 && hole.size(2) == 2,
                "hole must be a contiguous uint8 [H,S,2] tensor");
    TORCH_CHECK(board.scalar_type() == at::kByte && board.is_contiguous() && board.dim() == 2 && board.size(1) == 5 &&
                    board.size(0) == hole.size(0) && board.device() == hole.device(),
                "board must be a contiguous uint8 [H,5] tensor on the device of hole");
    const uint8_t *act = nullptr;
    at::Tensor act_u8;
    if (active.has_value()) {
        TORCH_CHECK(active->is_cuda() && active->device() == hole.device() && active->numel() == hole.size(0) * hole.size(1),
                    "active must be [H,S] on the device of hole");
        act_u8 = active->to(at::kByte).contiguous();
        act = act_u8.data_ptr<uint8_t>();
    }
    Enter e(hole);
    at::Tensor ranks = at::empty({hole.size(0), hole.size(1)}, hole.options().dtype(at::kShort));
    at::Tensor winners = at::empty({hole.size(0), 2}, hole.options().dtype(at::kInt));
    check(holdem_table_showdown(hole.data_ptr<uint8_t>(), board.data_ptr<uint8_t>(), act, hole.size(0), (int)hole.size(1),
                                (uint16_t *)ranks.data_ptr(), (uint32_t *)winners.data_ptr(), e.stream));
    return std::make_tuple(ranks, winners);
}

// Decision.decision_based_on_blind_position (Decision.py:4-48) straight from result rows; pos: uint8 [N] or None + pos_all.
at::Tensor decide_from_counts(const at::Tensor &rows, const c10::optional<at::Tensor> &pos, int64_t pos_all) {
    require(rows, "rows", at::kInt, HOLDEM_HP_ROW);
    const uint8_t *p = nullptr;
    if (pos.has_value()) {
        require(*pos, "pos", at::kByte, -1);
        TORCH_CHECK(pos->numel() == rows.size(0) && pos->device() == rows.device(), "pos must be uint8 [N] on the device of rows");
        p = pos->data_ptr<uint8_t>();
    }
    Enter e(rows);
    at::Tensor out = at::empty({rows.size(0)}, rows.options().dtype(at::kChar));
    check(holdem_decide_from_counts((const uint32_t *)rows.data_ptr(), p, (int)pos_all, rows.size(0),
                                    (int8_t *)out.data_ptr(), nullptr, e.stream));
    return out;
}

at::Tensor decide(const at::Tensor &hs_, const at::Tensor &ppot, const at::Tensor &npot,
                  const c10::optional<at::Tensor> &pos, int64_t pos_all) {
    require(hs_, "hand_strength", at::kDouble, -1);
    require(ppot, "ppot", at::kDouble, -1);
    require(npot, "npot", at::kDouble, -1);
    TORCH_CHECK(ppot.numel() == hs_.numel() && npot.numel() == hs_.numel(), "hand_strength, ppot and npot must have one shape");
    const uint8_t *p = nullptr;
    if (pos.has_value()) {
        require(*pos, "pos", at::kByte, -1);
        TORCH_CHECK(pos->numel() == hs_.numel() && pos->device() == hs_.device(), "pos must be uint8 [N] on the device of the inputs");
        p = pos->data_ptr<uint8_t>();
    }
    Enter e(hs_);
    at::Tensor out = at::empty(hs_.sizes(), hs_.options().dtype(at::kChar));
    check(holdem_decide_batch(hs_.data_ptr<double>(), ppot.data_ptr<double>(), npot.data_ptr<double>(), p, (int)pos_all,
                              hs_.numel(), (int8_t *)out.data_ptr(), e.stream));
    return out;
}

// utils.card_to_int wire ints (utils.py:82-85) -> situation rows.
at::Tensor pack_rank_major(const at::Tensor &hole, const at::Tensor &board) {
    require(hole, "hole", at::kLong, 2);
    require(board, "board", at::kLong, -1);
    TORCH_CHECK(board.dim() == 2 && board.size(0) == hole.size(0) && board.size(1) >= 3 && board.size(1) <= 5 &&
                    board.device() == hole.device(), "board must be int64 [N,3..5] on the device of hole");
    Enter e(hole);
    at::Tensor sit = at::empty({hole.size(0), 8}, hole.options().dtype(at::kByte));
    check(holdem_pack_rank_major(hole.data_ptr<int64_t>(), board.data_ptr<int64_t>(), (int)board.size(1), hole.size(0),
                                 sit.data_ptr<uint8_t>(), e.stream));
    return sit;
}

// Content hash of result rows (int32 [n, words]) -> int64 [1]; shards combine by (wrapping) addition.
at::Tensor rows_hash(const at::Tensor &rows, int64_t first_row) {
    require(rows, "rows", at::kInt, -1);
    TORCH_CHECK(rows.dim() == 2, "rows must be [n, words]");
    Enter e(rows);
    at::Tensor out = at::empty({1}, rows.options().dtype(at::kLong));
    check(holdem_rows_hash((const uint32_t *)rows.data_ptr(), rows.size(0), (int)rows.size(1), first_row,
                           (uint64_t *)out.data_ptr(), e.stream));
    return out;
}

}  // namespace

TORCH_LIBRARY(holdem, m) {
    m.def("eval(Tensor cards, int n_cards, bool strict=False) -> Tensor");
    m.def("category(Tensor cards, int n_cards) -> Tensor");
    m.def("hs(Tensor sit, int mode) -> Tensor");
    m.def("hp(Tensor sit, int mode) -> Tensor");
    m.def("hp_out(Tensor sit, int mode, Tensor(a!) out) -> ()");
    m.def("hp_peers(Tensor sit, int mode, Tensor(a!) out, int[] peer_ptrs) -> ()");
    m.def("table_hp(Tensor hole, Tensor board, int n_board, int mode) -> Tensor");
    m.def("table_showdown(Tensor hole, Tensor board, Tensor? active) -> (Tensor, Tensor)");
    m.def("decide(Tensor hand_strength, Tensor ppot, Tensor npot, Tensor? pos, int pos_all) -> Tensor");
    m.def("decide_from_counts(Tensor rows, Tensor? pos, int pos_all) -> Tensor");
    m.def("pack_rank_major(Tensor hole, Tensor board) -> Tensor");
    m.def("rows_hash(Tensor rows, int first_row) -> Tensor");
}

// CUDA only: there is deliberately no CPU kernel behind these schemas.
TORCH_LIBRARY_IMPL(holdem, CUDA, m) {
    m.impl("eval", eval);
    m.impl("category", category);
    m.impl("hs", hs);
    m.impl("hp", hp);
    m.impl("hp_out", hp_out);
    m.impl("hp_peers", hp_peers);
    m.impl("table_hp", table_hp);
    m.impl("table_showdown", table_showdown);
    m.impl("decide", decide);
    m.impl("decide_from_counts", decide_from_counts);
    m.impl("pack_rank_major", pack_rank_major);
    m.impl("rows_hash", rows_hash);
}
