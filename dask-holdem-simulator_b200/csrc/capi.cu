// capi.cu -- the C ABI of include/holdem_b200.h: table upload, argument checking, error reporting and the
// host-buffer entry points (pinned staging + chunked H2D / kernel / D2H pipeline).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/holdem_b200.h"
#include "device_common.cuh"
#include "launch.h"
#include "tables.h"

namespace holdem {
std::atomic<unsigned long long> g_kernel_launches{0};
}

namespace {

using namespace holdem;

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(HOLDEM_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));     \
    } while (0)

struct DeviceState {
    std::atomic<bool> ready{false};
    int sms = 0;
    uint8_t *image = nullptr;
    DevTables T;
    // pinned staging + device scratch for the *_host entry points (grown on demand, guarded by host_mu)
    std::mutex host_mu;
    uint8_t *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    uint8_t *d_buf = nullptr;
    size_t d_buf_bytes = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    // Ring of 16-byte device slots (flag word + work counters) handed to the HP dispatch, one per call.  A slot is reused
    // only behind the event recorded at the end of the call that used it last (stream-ordered wait, no host block), so any
    // number of calls may be in flight on any number of streams.
    int *flags = nullptr;
    cudaEvent_t slot_done[1024] = {};
    bool slot_used[1024] = {};
    unsigned flag_next = 0;
    std::mutex flag_mu;
    cudaMemPool_t pool = nullptr;   // stream-ordered scratch (whole-table entry point); keeps its memory between calls
};
constexpr unsigned kFlagRing = 1024;

struct FlagSlot {           // RAII: acquire() under the device's flag mutex, release() records the slot's event on the stream
    DeviceState *S = nullptr;
    unsigned idx = 0;
    cudaStream_t st = nullptr;
    int *ptr = nullptr;
    std::unique_lock<std::mutex> lock;
    cudaError_t acquire(DeviceState &state, cudaStream_t stream) {
        S = &state;
        st = stream;
        lock = std::unique_lock<std::mutex>(state.flag_mu);
        idx = state.flag_next++ % kFlagRing;
        ptr = state.flags + 4 * idx;
        if (state.slot_used[idx]) return cudaStreamWaitEvent(stream, state.slot_done[idx], 0);
        cudaError_t e = cudaEventCreateWithFlags(&state.slot_done[idx], cudaEventDisableTiming);
        if (e == cudaSuccess) state.slot_used[idx] = true;
        return e;
    }
    cudaError_t release() {
        cudaError_t e = S ? cudaEventRecord(S->slot_done[idx], st) : cudaSuccess;
        S = nullptr;
        if (lock.owns_lock()) lock.unlock();
        return e;
    }
    ~FlagSlot() { if (S) release(); }
};


constexpr int kMaxDevices = 64;
DeviceState g_dev[kMaxDevices];
std::mutex g_init_mu;

int current_state(DeviceState **out) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
        return fail(HOLDEM_ERR_NO_DEVICE, "no current CUDA device");
    if (!g_dev[dev].ready) return fail(HOLDEM_ERR_NOT_INIT, "holdem_init(%d) has not been called", dev);
    *out = &g_dev[dev];
    return HOLDEM_OK;
}

int ensure_host_buffers(DeviceState &S, size_t pin_bytes, size_t dev_bytes) {
    if (S.h_pin_bytes < pin_bytes) {
        if (S.h_pin) cudaFreeHost(S.h_pin);
        S.h_pin = nullptr;
        S.h_pin_bytes = 0;
        CUDA_TRY(cudaMallocHost(&S.h_pin, pin_bytes));
        S.h_pin_bytes = pin_bytes;
    }
    if (S.d_buf_bytes < dev_bytes) {
        if (S.d_buf) cudaFree(S.d_buf);
        S.d_buf = nullptr;
        S.d_buf_bytes = 0;
        CUDA_TRY(cudaMalloc(&S.d_buf, dev_bytes));
        S.d_buf_bytes = dev_bytes;
    }
    for (auto &st : S.streams)
        if (!st) CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    return HOLDEM_OK;
}

// Generic chunked host pipeline: rows of in_row bytes -> rows of out_row bytes, double-buffered over two streams so
// that the H2D copy of chunk k+1 overlaps the kernel of chunk k and the D2H copy of chunk k-1.
template <class Launch>
int host_pipeline(const void *h_in, size_t in_row, void *h_out, size_t out_row, int64_t n, int64_t chunk_rows,
                  Launch launch) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n == 0) return HOLDEM_OK;
    std::lock_guard<std::mutex> lock(S->host_mu);
    if (chunk_rows > n) chunk_rows = n;
    const size_t in_chunk = (size_t)chunk_rows * in_row, out_chunk = (size_t)chunk_rows * out_row;
    const size_t slot = ((in_chunk + out_chunk + 255) / 256) * 256;
    // Caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) are used directly:
    // no staging copies on the host side.
    auto is_pinned = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool direct = is_pinned(h_in) && is_pinned(h_out);
    rc = ensure_host_buffers(*S, direct ? 0 : 2 * slot, 2 * slot);
    if (rc) return rc;
    if (direct) {
        int64_t done = 0;
        for (int k = 0; done < n; ++k) {
            const int b = k & 1;
            const int64_t rows = n - done < chunk_rows ? n - done : chunk_rows;
            uint8_t *d_in = S->d_buf + b * slot, *d_out = d_in + in_chunk;
            if (k >= 2) CUDA_TRY(cudaStreamSynchronize(S->streams[b]));   // the device slot is free again
            CUDA_TRY(cudaMemcpyAsync(d_in, (const uint8_t *)h_in + (size_t)done * in_row, (size_t)rows * in_row,
                                     cudaMemcpyHostToDevice, S->streams[b]));
            CUDA_TRY(launch(*S, d_in, rows, d_out, S->streams[b]));
            CUDA_TRY(cudaMemcpyAsync((uint8_t *)h_out + (size_t)done * out_row, d_out, (size_t)rows * out_row,
                                     cudaMemcpyDeviceToHost, S->streams[b]));
            done += rows;
        }
        CUDA_TRY(cudaStreamSynchronize(S->streams[0]));
        CUDA_TRY(cudaStreamSynchronize(S->streams[1]));
        return HOLDEM_OK;
    }
    int64_t done = 0;
    int64_t pending_rows[2] = {0, 0};
    int64_t pending_off[2] = {0, 0};
    int k = 0;
    auto drain = [&](int b) -> int {
        if (pending_rows[b] == 0) return HOLDEM_OK;
        CUDA_TRY(cudaStreamSynchronize(S->streams[b]));
        std::memcpy((uint8_t *)h_out + (size_t)pending_off[b] * out_row, S->h_pin + b * slot + in_chunk,
                    (size_t)pending_rows[b] * out_row);
        pending_rows[b] = 0;
        return HOLDEM_OK;
    };
    while (done < n) {
        int b = k & 1;
        rc = drain(b);
        if (rc) return rc;
        int64_t rows = n - done < chunk_rows ? n - done : chunk_rows;
        uint8_t *hp_in = S->h_pin + b * slot, *hp_out = hp_in + in_chunk;
        uint8_t *d_in = S->d_buf + b * slot, *d_out = d_in + in_chunk;
        std::memcpy(hp_in, (const uint8_t *)h_in + (size_t)done * in_row, (size_t)rows * in_row);
        CUDA_TRY(cudaMemcpyAsync(d_in, hp_in, (size_t)rows * in_row, cudaMemcpyHostToDevice, S->streams[b]));
        CUDA_TRY(launch(*S, d_in, rows, d_out, S->streams[b]));
        CUDA_TRY(cudaMemcpyAsync(hp_out, d_out, (size_t)rows * out_row, cudaMemcpyDeviceToHost, S->streams[b]));
        pending_rows[b] = rows;
        pending_off[b] = done;
        done += rows;
        ++k;
    }
    rc = drain(0);
    if (rc) return rc;
    return drain(1);
}

}  // namespace

extern "C" {

int holdem_abi_version(void) { return HOLDEM_ABI_VERSION; }

const char *holdem_last_error(void) { return g_err; }

int holdem_selftest_tables(void) {
    char msg[256];
    if (holdem::selftest(msg, sizeof msg)) return fail(HOLDEM_ERR_INTERNAL, "table self-check failed: %s", msg);
    return HOLDEM_OK;
}

int holdem_init(int device) {
    if (device < 0 || device >= kMaxDevices) return fail(HOLDEM_ERR_BAD_ARG, "device index %d out of range", device);
    if (g_dev[device].ready.load(std::memory_order_acquire)) {     // fast path: no lock, no device-count query
        CUDA_TRY(cudaSetDevice(device));
        return HOLDEM_OK;
    }
    std::lock_guard<std::mutex> lock(g_init_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(HOLDEM_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device >= count) return fail(HOLDEM_ERR_NO_DEVICE, "device %d requested but only %d present", device, count);
    DeviceState &S = g_dev[device];
    CUDA_TRY(cudaSetDevice(device));
    if (S.ready) return HOLDEM_OK;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(HOLDEM_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    char msg[256];
    if (holdem::selftest(msg, sizeof msg)) return fail(HOLDEM_ERR_INTERNAL, "table self-check failed: %s", msg);
    const Tables &t = get_tables();
    TableImage im = make_image(t);
    CUDA_TRY(cudaMalloc(&S.image, im.bytes.size()));
    CUDA_TRY(cudaMemcpy(S.image, im.bytes.data(), im.bytes.size(), cudaMemcpyHostToDevice));
    S.T.flush = reinterpret_cast<const uint16_t *>(S.image + im.off_flush);
    S.T.pairs = reinterpret_cast<const uint16_t *>(S.image + im.off_pairs);
    S.T.flop_task_hdr = reinterpret_cast<const uint32_t *>(S.image + im.off_task_hdr);
    S.T.flop_lanes = reinterpret_cast<const uint32_t *>(S.image + im.off_lanes);
    S.T.nf4_ranks = reinterpret_cast<const uint16_t *>(S.image + im.off_nf4);
    S.T.board_t = reinterpret_cast<const uint16_t *>(S.image + im.off_board_t);
    S.T.board_opp5 = reinterpret_cast<const uint16_t *>(S.image + im.off_board_opp5);
    S.T.own_our7 = reinterpret_cast<const uint16_t *>(S.image + im.off_own_our7);
    S.T.own_rank5 = reinterpret_cast<const uint16_t *>(S.image + im.off_own_rank5);
    S.T.board4_t = reinterpret_cast<const uint16_t *>(S.image + im.off_board4_t);
    S.T.board4_opp6 = reinterpret_cast<const uint16_t *>(S.image + im.off_board4_opp6);
    S.T.own6_our7 = reinterpret_cast<const uint16_t *>(S.image + im.off_own6_our7);
    S.T.own6_rank6 = reinterpret_cast<const uint16_t *>(S.image + im.off_own6_rank6);
    S.T.n_flop_tasks = im.n_tasks;
    S.T.turn_task_hdr = reinterpret_cast<const uint32_t *>(S.image + im.off_turn_hdr);
    S.T.turn_lanes = reinterpret_cast<const uint32_t *>(S.image + im.off_turn_lanes);
    S.T.nf3_ranks = reinterpret_cast<const uint16_t *>(S.image + im.off_nf3);
    S.T.n_turn_tasks = im.n_turn_tasks;
    for (int q = 0; q < 2; ++q) {
        S.T.ref_task_hdr[q] = reinterpret_cast<const uint32_t *>(S.image + im.off_ref_hdr[q]);
        S.T.ref_lanes[q] = reinterpret_cast<const uint32_t *>(S.image + im.off_ref_lanes[q]);
        S.T.n_ref_tasks[q] = im.n_ref_tasks[q];
    }
    for (int i = 0; i < 3; ++i) {
        NfHash &h = S.T.nf[i];
        h.disp = reinterpret_cast<const uint16_t *>(S.image + im.off_disp[i]);
        h.val = reinterpret_cast<const uint16_t *>(S.image + im.off_val[i]);
        h.mul_b = t.nf[i].mul_b;
        h.mul_s = t.nf[i].mul_s;
        h.bshift = 32 - t.nf[i].bucket_bits;
        h.sshift = 32 - t.nf[i].slot_bits;
        h.smask = (1u << t.nf[i].slot_bits) - 1u;
        S.T.disp_entries[i] = (uint32_t)t.nf[i].disp.size();
        S.T.val_entries[i] = (uint32_t)t.nf[i].val.size();
    }
    for (int k = 3; k <= 5; ++k) {          // reference-mode category tables: built on the device (71 MB in all, ~1 ms)
        uint8_t *tab = nullptr;
        CUDA_TRY(cudaMalloc(&tab, ref_t_table_bytes(k)));
        CUDA_TRY(launch_ref_t_table(tab, k, prop.multiProcessorCount, nullptr));
        S.T.ref_t[k - 3] = tab;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMalloc(&S.flags, 4 * kFlagRing * sizeof(int)));     // 16-byte slots: flag word + work counters
    CUDA_TRY(cudaMemset(S.flags, 0, 4 * kFlagRing * sizeof(int)));
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CUDA_TRY(cudaMemPoolCreate(&S.pool, &props));
        unsigned long long keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(S.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    S.sms = prop.multiProcessorCount;
    S.ready.store(true, std::memory_order_release);
    return HOLDEM_OK;
}

int holdem_sm_count(int device) {
    if (device < 0 || device >= kMaxDevices || !g_dev[device].ready)
        return fail(HOLDEM_ERR_NOT_INIT, "holdem_init(%d) has not been called", device);
    return g_dev[device].sms;
}

// ---- evaluator ---------------------------------------------------------------------------------------------------
int holdem_eval_batch(const uint8_t *d_cards, int n_cards, int64_t n, uint16_t *d_rank_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_cards || !d_rank_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (n_cards < 5 || n_cards > 7)
        return fail(HOLDEM_ERR_BAD_ARG, "n_cards must be 5, 6 or 7 (got %d): Evaluator.hand_size_map", n_cards);
    CUDA_TRY(launch_eval_batch(S->T, d_cards, n_cards, n, d_rank_out, S->sms, false, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_eval_batch_strict(const uint8_t *d_cards, int n_cards, int64_t n, uint16_t *d_rank_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_cards || !d_rank_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (n_cards < 5 || n_cards > 7)
        return fail(HOLDEM_ERR_BAD_ARG, "n_cards must be 5, 6 or 7 (got %d): Evaluator.hand_size_map", n_cards);
    CUDA_TRY(launch_eval_batch(S->T, d_cards, n_cards, n, d_rank_out, S->sms, true, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_eval_batch_host(const uint8_t *h_cards, int n_cards, int64_t n, uint16_t *h_rank_out) {
    if (n < 0 || (n > 0 && (!h_cards || !h_rank_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (n_cards < 5 || n_cards > 7) return fail(HOLDEM_ERR_BAD_ARG, "n_cards must be 5, 6 or 7 (got %d)", n_cards);
    int rc = host_pipeline(h_cards, 8, h_rank_out, 2, n, 1 << 22,
                           [&](DeviceState &S, uint8_t *d_in, int64_t rows, uint8_t *d_out, cudaStream_t st) {
                               return launch_eval_batch(S.T, d_in, n_cards, rows, (uint16_t *)d_out, S.sms, true, st);
                           });
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i)
        if (h_rank_out[i] == 0) return fail(HOLDEM_ERR_BAD_INPUT, "row %lld is malformed", (long long)i);
    return HOLDEM_OK;
}

int holdem_exhaustive7(uint64_t *d_hist, uint64_t *d_sums, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (!d_hist || !d_sums) return fail(HOLDEM_ERR_BAD_ARG, "null pointer");
    CUDA_TRY(launch_exhaustive7(S->T, d_hist, d_sums, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_exhaustive7_host(uint64_t *h_hist, uint64_t *h_sums) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (!h_hist || !h_sums) return fail(HOLDEM_ERR_BAD_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(S->host_mu);
    rc = ensure_host_buffers(*S, 256, 7465 * 8);
    if (rc) return rc;
    uint64_t *d = reinterpret_cast<uint64_t *>(S->d_buf);
    CUDA_TRY(launch_exhaustive7(S->T, d, d + 7463, S->sms, S->streams[0]));
    CUDA_TRY(cudaMemcpyAsync(h_hist, d, 7463 * 8, cudaMemcpyDeviceToHost, S->streams[0]));
    CUDA_TRY(cudaMemcpyAsync(h_sums, d + 7463, 2 * 8, cudaMemcpyDeviceToHost, S->streams[0]));
    CUDA_TRY(cudaStreamSynchronize(S->streams[0]));
    return HOLDEM_OK;
}

// ---- literal ranker ------------------------------------------------------------------------------------------------
int holdem_category_batch(const uint8_t *d_cards, int n_cards, int row_stride, int64_t n, uint8_t *d_cat_out,
                          void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_cards || !d_cat_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (n_cards < 2 || n_cards > 9 || row_stride < n_cards)
        return fail(HOLDEM_ERR_BAD_ARG, "n_cards must be 2..9 and row_stride >= n_cards (got %d, %d)", n_cards, row_stride);
    CUDA_TRY(launch_category_batch(d_cards, n_cards, row_stride, n, d_cat_out, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_category_batch_host(const uint8_t *h_cards, int n_cards, int row_stride, int64_t n, uint8_t *h_cat_out) {
    if (n < 0 || (n > 0 && (!h_cards || !h_cat_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (n_cards < 2 || n_cards > 9 || row_stride < n_cards)
        return fail(HOLDEM_ERR_BAD_ARG, "n_cards must be 2..9 and row_stride >= n_cards (got %d, %d)", n_cards, row_stride);
    int rc = host_pipeline(h_cards, (size_t)row_stride, h_cat_out, 1, n, 1 << 22,
                           [&](DeviceState &S, uint8_t *d_in, int64_t rows, uint8_t *d_out, cudaStream_t st) {
                               return launch_category_batch(d_in, n_cards, row_stride, rows, d_out, S.sms, st);
                           });
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i)
        if (h_cat_out[i] == 0xFF) return fail(HOLDEM_ERR_BAD_INPUT, "row %lld is malformed", (long long)i);
    return HOLDEM_OK;
}

// ---- HS / HP ---------------------------------------------------------------------------------------------------------
static int check_mode(int mode) {
    if (mode != HOLDEM_MODE_TREYS && mode != HOLDEM_MODE_REFERENCE)
        return fail(HOLDEM_ERR_BAD_ARG, "unknown mode %d", mode);
    return HOLDEM_OK;
}

int holdem_hs_batch(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if ((rc = check_mode(mode))) return rc;
    CUDA_TRY(launch_hs_batch(S->T, d_sit, n, mode, d_out, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_hs_batch_host(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out) {
    if (n < 0 || (n > 0 && (!h_sit || !h_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    int rc = check_mode(mode);
    if (rc) return rc;
    rc = host_pipeline(h_sit, 8, h_out, 12, n, 1 << 20,
                       [&](DeviceState &S, uint8_t *d_in, int64_t rows, uint8_t *d_out, cudaStream_t st) {
                           return launch_hs_batch(S.T, d_in, rows, mode, (uint32_t *)d_out, S.sms, st);
                       });
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i)
        if (h_out[i * 3] == HOLDEM_BAD_ROW) return fail(HOLDEM_ERR_BAD_INPUT, "situation %lld is malformed", (long long)i);
    return HOLDEM_OK;
}

// Which implementation serves each street of a Treys-mode HP call.  Production = {4, 4, 0}; the others are reached through the
// explicit *_variant entry points only (A/B measurements and independent cross-checks).
struct HpImpl {
    int flop = 4;      // 4 rank-multiset kernel, 5 the same without the closed-form add path, 3 / 2 the 4-set kernels
    int turn = 4;      // 4 rank-multiset kernel, 3 the 3-set walk
    int river = 0;     // 0 rank-pair HS kernel in its HP-row form, 1 plain enumeration
    bool order = true; // work the flop rows heaviest board texture first (false: row order; holdem_hp_flop_variant 6)
    int ref = 5;       // HOLDEM_MODE_REFERENCE: 5 rank-multiset kernel, 4 the earlier 4-set walk + arithmetic irregular runouts
};

// Treys mode: the flop kernel takes every 3-card-board row and flags what else it saw; turn rows go to the turn kernel, river
// rows (HS only) to the rank-pair kernel.  Both exit at once when not needed.  Reference mode: hp_ref_kernel, all streets.
static cudaError_t hp_dispatch(DeviceState &S, const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out,
                               cudaStream_t st, const PeerOut &peers = PeerOut{}, int **flag_out = nullptr,
                               const HpImpl &impl = HpImpl{}) {
    // flag word of this call: bit 0 / 1 = the flop kernel met turn / river rows, bit 2 = some kernel marked a malformed row
    // (flag_out receives its address when every kernel of the call maintains bit 2, else nullptr)
    FlagSlot slot;
    cudaError_t e = slot.acquire(S, st);
    if (e != cudaSuccess) return e;
    int *flag = slot.ptr;
    if (mode == HOLDEM_MODE_REFERENCE && impl.ref == 5) {
        e = launch_hp_ref5(S.T, d_sit, n, d_out, S.sms, flag, peers, st);      // stores its rows into the peers itself
        if (e != cudaSuccess) return e;
        if (flag_out) *flag_out = flag;
        return slot.release();
    }
    if (mode == HOLDEM_MODE_REFERENCE) {
        e = cudaMemsetAsync(flag, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
        e = launch_hp_ref(S.T, d_sit, n, d_out, S.sms, flag, st);
        if (e != cudaSuccess) return e;
        if (flag_out) *flag_out = flag;
        e = launch_rows_to_peers(d_sit, n, d_out, peers, S.sms, nullptr, st);
        if (e != cudaSuccess) return e;
        return slot.release();
    }
    const bool old_flop = impl.flop == 2 || impl.flop == 3, old_turn = impl.turn == 3;
    const bool fused = !old_flop && !old_turn;          // the production kernels store their rows into the peers themselves
    const PeerOut none{};
    // batches worth ordering by board texture get a stream-ordered scratch buffer for the work order (private pool)
    uint32_t *order = nullptr;
    if (fused && impl.order && n >= 4096) {
        e = cudaMallocFromPoolAsync((void **)&order, street_order_scratch_bytes(n), S.pool, st);
        if (e != cudaSuccess) return e;
        e = launch_street_order(d_sit, n, order, S.sms, st);
        if (e != cudaSuccess) { cudaFreeAsync(order, st); return e; }
    }
    e = old_flop ? launch_hp_flop_treys(S.T, d_sit, n, d_out, S.sms, flag, impl.flop, st)
                 : launch_hp_flop_treys_v4(S.T, d_sit, n, d_out, S.sms, flag, nullptr, fused ? peers : none, st, impl.flop != 5, order,
                                           impl.flop != 7);
    if (e == cudaSuccess)
        e = old_turn ? launch_hp_turn_treys(S.T, d_sit, n, d_out, S.sms, flag, st)
                     : launch_hp_turn_treys_v4(S.T, d_sit, n, d_out, S.sms, flag, fused ? peers : none, st, order);
    if (order != nullptr) cudaFreeAsync(order, st);
    if (e != cudaSuccess) return e;
    e = impl.river == 1 ? launch_hp_direct(S.T, d_sit, n, mode, d_out, S.sms, flag, 1, st)
                        : launch_hp_river_rows(S.T, d_sit, n, d_out, S.sms, flag, st);
    if (e != cudaSuccess) return e;
    if (flag_out) *flag_out = fused ? flag : nullptr;
    // river rows only (exits at once when there are none); every row when an earlier kernel was selected
    e = launch_rows_to_peers(d_sit, n, d_out, peers, S.sms, fused ? flag : nullptr, st);
    if (e != cudaSuccess) return e;
    return slot.release();
}

int holdem_hp_batch(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if ((rc = check_mode(mode))) return rc;
    CUDA_TRY(hp_dispatch(*S, d_sit, n, mode, d_out, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_hp_batch_peers(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *const *peer_out, int n_peers,
                          void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !peer_out)) return fail(HOLDEM_ERR_BAD_ARG, "0..8 peer buffers");
    if ((rc = check_mode(mode))) return rc;
    PeerOut peers{};
    peers.n = n_peers;
    for (int k = 0; k < n_peers; ++k) {
        if (!peer_out[k]) return fail(HOLDEM_ERR_BAD_ARG, "null peer buffer %d", k);
        peers.p[k] = (uint32_t *)peer_out[k];
    }
    CUDA_TRY(hp_dispatch(*S, d_sit, n, mode, d_out, (cudaStream_t)stream, peers));
    return HOLDEM_OK;
}

int holdem_hp_batch_direct(const uint8_t *d_sit, int64_t n, int mode, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if ((rc = check_mode(mode))) return rc;
    CUDA_TRY(launch_hp_direct(S->T, d_sit, n, mode, d_out, S->sms, nullptr, 0, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_hp_flop_variant(const uint8_t *d_sit, int64_t n, int variant, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (variant < 2 || variant > 7) return fail(HOLDEM_ERR_BAD_ARG, "unknown flop kernel variant %d", variant);
    FlagSlot slot;
    CUDA_TRY(slot.acquire(*S, (cudaStream_t)stream));
    if (variant >= 4) {
        uint32_t *order = nullptr;
        if (variant != 6 && n >= 4096) {
            CUDA_TRY(cudaMallocFromPoolAsync((void **)&order, street_order_scratch_bytes(n), S->pool, (cudaStream_t)stream));
            cudaError_t eo = launch_street_order(d_sit, n, order, S->sms, (cudaStream_t)stream);
            if (eo != cudaSuccess) { cudaFreeAsync(order, (cudaStream_t)stream); CUDA_TRY(eo); }
        }
        cudaError_t e = launch_hp_flop_treys_v4(S->T, d_sit, n, d_out, S->sms, slot.ptr, nullptr, PeerOut{}, (cudaStream_t)stream,
                                                variant != 5, order, variant != 7);
        if (order != nullptr) cudaFreeAsync(order, (cudaStream_t)stream);
        CUDA_TRY(e);
    } else CUDA_TRY(launch_hp_flop_treys(S->T, d_sit, n, d_out, S->sms, slot.ptr, variant, (cudaStream_t)stream));
    CUDA_TRY(slot.release());
    return HOLDEM_OK;
}

int holdem_hp_turn_variant(const uint8_t *d_sit, int64_t n, int variant, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (variant != 3 && variant != 4 && variant != 6) return fail(HOLDEM_ERR_BAD_ARG, "unknown turn kernel variant %d", variant);
    if (variant == 3) CUDA_TRY(launch_hp_turn_treys(S->T, d_sit, n, d_out, S->sms, nullptr, (cudaStream_t)stream));
    else {
        // the kernel's work counter lives in a flag slot (word 2); word 0 stays 0 and is not consulted (run_flag == nullptr)
        FlagSlot slot;
        CUDA_TRY(slot.acquire(*S, (cudaStream_t)stream));
        CUDA_TRY(cudaMemsetAsync(slot.ptr, 0, 4 * sizeof(int), (cudaStream_t)stream));
        uint32_t *order = nullptr;
        if (variant == 4 && n >= 4096) {
            CUDA_TRY(cudaMallocFromPoolAsync((void **)&order, street_order_scratch_bytes(n), S->pool, (cudaStream_t)stream));
            cudaError_t eo = launch_street_order(d_sit, n, order, S->sms, (cudaStream_t)stream);
            if (eo != cudaSuccess) { cudaFreeAsync(order, (cudaStream_t)stream); CUDA_TRY(eo); }
        }
        cudaError_t e = launch_hp_turn_treys_v4_counter(S->T, d_sit, n, d_out, S->sms, slot.ptr, PeerOut{}, (cudaStream_t)stream, order);
        if (order != nullptr) cudaFreeAsync(order, (cudaStream_t)stream);
        CUDA_TRY(e);
        CUDA_TRY(slot.release());
    }
    return HOLDEM_OK;
}

int holdem_hp_flop_phase_cycles(const uint8_t *d_sit, int64_t n, uint32_t *d_out, uint64_t *d_cycles, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || !d_cycles || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    FlagSlot slot;
    CUDA_TRY(slot.acquire(*S, (cudaStream_t)stream));
    CUDA_TRY(cudaMemsetAsync(d_cycles, 0, 8 * sizeof(uint64_t), (cudaStream_t)stream));
    CUDA_TRY(launch_hp_flop_treys_v4(S->T, d_sit, n, d_out, S->sms, slot.ptr, (unsigned long long *)d_cycles, PeerOut{},
                                     (cudaStream_t)stream));
    CUDA_TRY(slot.release());
    return HOLDEM_OK;
}

static int hp_batch_host_impl(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out, bool zero_copy) {
    if (n < 0 || (n > 0 && (!h_sit || !h_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    int rc = check_mode(mode);
    if (rc) return rc;
    // Page-locked caller buffers: one H2D copy of the 8-byte rows, ONE launch over the whole batch, and the kernels store
    // their 64-byte result rows straight into the caller's (mapped) host buffer -- the device-to-host transfer runs under
    // the computation instead of after it.  holdem_hp_batch_host_staged selects the chunked copy pipeline instead.
    bool done = false;
    if (zero_copy && n > 0) {
        DeviceState *S = nullptr;
        if ((rc = current_state(&S))) return rc;
        cudaPointerAttributes ai, ao;
        const bool pinned = cudaPointerGetAttributes(&ai, h_sit) == cudaSuccess && ai.type == cudaMemoryTypeHost &&
                            cudaPointerGetAttributes(&ao, h_out) == cudaSuccess && ao.type == cudaMemoryTypeHost &&
                            ao.devicePointer != nullptr;
        cudaGetLastError();
        if (pinned) {
            std::lock_guard<std::mutex> lock(S->host_mu);
            if ((rc = ensure_host_buffers(*S, 0, (size_t)n * 8))) return rc;
            cudaStream_t st = S->streams[0];
            CUDA_TRY(cudaMemcpyAsync(S->d_buf, h_sit, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            int *d_flag = nullptr;
            int h_flag = 4;
            CUDA_TRY(hp_dispatch(*S, S->d_buf, n, mode, (uint32_t *)ao.devicePointer, st, PeerOut{}, &d_flag));
            if (d_flag) CUDA_TRY(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            done = true;
            if (!(h_flag & 4)) return HOLDEM_OK;      // no kernel marked a malformed row: nothing to look for in the results
        }
    }
    if (!done) {
        rc = host_pipeline(h_sit, 8, h_out, 64, n, 1 << 14,
                           [&](DeviceState &S, uint8_t *d_in, int64_t rows, uint8_t *d_out, cudaStream_t st) {
                               return hp_dispatch(S, d_in, rows, mode, (uint32_t *)d_out, st);
                           });
        if (rc) return rc;
    }
    for (int64_t i = 0; i < n; ++i)
        if (h_out[i * 16 + 15] == HOLDEM_BAD_ROW) return fail(HOLDEM_ERR_BAD_INPUT, "situation %lld is malformed", (long long)i);
    return HOLDEM_OK;
}

int holdem_hp_batch_host(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out) {
    return hp_batch_host_impl(h_sit, n, mode, h_out, true);
}

int holdem_hp_batch_host_staged(const uint8_t *h_sit, int64_t n, int mode, uint32_t *h_out) {
    return hp_batch_host_impl(h_sit, n, mode, h_out, false);
}

int holdem_hp_batch_variant(const uint8_t *d_sit, int64_t n, int flop_variant, int turn_variant, int river_variant,
                            uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (flop_variant < 2 || flop_variant > 7 || (turn_variant != 3 && turn_variant != 4) || river_variant < 0 || river_variant > 1)
        return fail(HOLDEM_ERR_BAD_ARG, "unknown kernel variant (flop 2..7, turn 3..4, river 0..1)");
    HpImpl impl;
    impl.flop = flop_variant == 6 ? 4 : flop_variant; impl.order = flop_variant != 6; impl.turn = turn_variant; impl.river = river_variant;
    CUDA_TRY(hp_dispatch(*S, d_sit, n, HOLDEM_MODE_TREYS, d_out, (cudaStream_t)stream, PeerOut{}, nullptr, impl));
    return HOLDEM_OK;
}

int holdem_hp_ref_variant(const uint8_t *d_sit, int64_t n, int variant, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (variant != 4 && variant != 5) return fail(HOLDEM_ERR_BAD_ARG, "unknown reference-mode kernel variant %d", variant);
    HpImpl impl;
    impl.ref = variant;
    CUDA_TRY(hp_dispatch(*S, d_sit, n, HOLDEM_MODE_REFERENCE, d_out, (cudaStream_t)stream, PeerOut{}, nullptr, impl));
    return HOLDEM_OK;
}

int holdem_hs_batch_variant(const uint8_t *d_sit, int64_t n, int mode, int variant, uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_sit || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if ((rc = check_mode(mode))) return rc;
    if (variant < 0 || variant > 1) return fail(HOLDEM_ERR_BAD_ARG, "unknown HS kernel variant %d", variant);
    CUDA_TRY(launch_hs_batch(S->T, d_sit, n, mode, d_out, S->sms, (cudaStream_t)stream, variant));
    return HOLDEM_OK;
}

int holdem_exhaustive7_variant(int variant, uint64_t *d_hist, uint64_t *d_sums, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (!d_hist || !d_sums) return fail(HOLDEM_ERR_BAD_ARG, "null pointer");
    if (variant < 0 || variant > 1) return fail(HOLDEM_ERR_BAD_ARG, "unknown sweep kernel variant %d", variant);
    CUDA_TRY(launch_exhaustive7(S->T, d_hist, d_sums, S->sms, (cudaStream_t)stream, variant));
    return HOLDEM_OK;
}

// ---- board-shared table evaluation ------------------------------------------------------------------------------------------
int holdem_table_hp_batch(const uint8_t *d_hole, const uint8_t *d_board, int64_t hands, int seats, int n_board, int mode,
                          uint32_t *d_out, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (hands < 0 || (hands > 0 && (!d_hole || !d_board || !d_out))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative size");
    if (seats < 1 || seats > 22) return fail(HOLDEM_ERR_BAD_ARG, "seats must be 1..22 (GameBoard.MAX_PLAYERS = 22), got %d", seats);
    if (n_board < 3 || n_board > 5) return fail(HOLDEM_ERR_BAD_ARG, "n_board must be 3, 4 or 5 (got %d)", n_board);
    if ((rc = check_mode(mode))) return rc;
    if (hands == 0) return HOLDEM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = hands * seats;
    uint8_t *d_sit = nullptr;
    CUDA_TRY(cudaMallocFromPoolAsync(&d_sit, (size_t)n * 8, S->pool, st));           // stream-ordered scratch: the expanded situation rows
    cudaError_t e = launch_table_rows(d_hole, d_board, hands, seats, n_board, d_sit, S->sms, st);
    if (e == cudaSuccess) e = hp_dispatch(*S, d_sit, n, mode, d_out, st);
    cudaFreeAsync(d_sit, st);
    CUDA_TRY(e);
    return HOLDEM_OK;
}

int holdem_table_showdown(const uint8_t *d_hole, const uint8_t *d_board, const uint8_t *d_active, int64_t hands, int seats,
                          uint16_t *d_ranks, uint32_t *d_winners, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (hands < 0 || (hands > 0 && (!d_hole || !d_board || !d_ranks || !d_winners)))
        return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative size");
    if (seats < 1 || seats > 22) return fail(HOLDEM_ERR_BAD_ARG, "seats must be 1..22 (GameBoard.MAX_PLAYERS = 22), got %d", seats);
    CUDA_TRY(launch_table_showdown(S->T, d_hole, d_board, d_active, hands, seats, d_ranks, d_winners, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

// ---- result-row hash, decision tree, wire-format packing (kernels_aux.cu) ------------------------------------------------
unsigned long long holdem_kernel_launches(void) { return holdem::g_kernel_launches.load(std::memory_order_relaxed); }

int holdem_rows_hash(const uint32_t *d_rows, int64_t n, int words_per_row, int64_t first_row, uint64_t *d_hash, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || !d_hash || words_per_row < 1 || (n > 0 && !d_rows)) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or bad size");
    CUDA_TRY(launch_rows_hash(d_rows, n, words_per_row, first_row, d_hash, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_decide_batch(const double *d_hs, const double *d_ppot, const double *d_npot, const uint8_t *d_pos, int pos_all,
                        int64_t n, int8_t *d_action, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_hs || !d_ppot || !d_npot || !d_action))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (!d_pos && (pos_all < 0 || pos_all > 2))
        return fail(HOLDEM_ERR_BAD_ARG, "Invalid position. Expected 'SB', 'BB', or 'NONE'.");   /* Decision.py:8-9 */
    CUDA_TRY(launch_decide(d_hs, d_ppot, d_npot, d_pos, pos_all, n, d_action, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_decide_from_counts(const uint32_t *d_rows, const uint8_t *d_pos, int pos_all, int64_t n, int8_t *d_action,
                              double *d_hs_ppot_npot, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_rows || !d_action))) return fail(HOLDEM_ERR_BAD_ARG, "null pointer or negative n");
    if (!d_pos && (pos_all < 0 || pos_all > 2))
        return fail(HOLDEM_ERR_BAD_ARG, "Invalid position. Expected 'SB', 'BB', or 'NONE'.");
    CUDA_TRY(launch_decide_from_counts(d_rows, d_pos, pos_all, n, d_action, d_hs_ppot_npot, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

int holdem_pack_rank_major(const int64_t *d_hole, const int64_t *d_board, int n_board, int64_t n, uint8_t *d_sit, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (n < 0 || n_board < 3 || n_board > 5 || (n > 0 && (!d_hole || !d_board || !d_sit)))
        return fail(HOLDEM_ERR_BAD_ARG, "null pointer, negative n or n_board outside 3..5");
    CUDA_TRY(launch_pack_rank_major(d_hole, d_board, n_board, n, d_sit, S->sms, (cudaStream_t)stream));
    return HOLDEM_OK;
}

// ---- micro-benchmark ---------------------------------------------------------------------------------------------------
int holdem_microbench_smem_gather(int table_bytes, int iters, int pattern, int blocks_per_sm, int threads,
                                  float *elapsed_ms, double *lookups, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (table_bytes < 1024 || table_bytes > 200 * 1024 || (table_bytes & (table_bytes - 1)) || iters < 1 ||
        threads < 32 || threads > 1024 || (threads & 31) || blocks_per_sm < 1 || !elapsed_ms || !lookups)
        return fail(HOLDEM_ERR_BAD_ARG, "bad micro-benchmark arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int blocks = S->sms * blocks_per_sm;
    CUDA_TRY(launch_smem_gather(table_bytes, iters, pattern, blocks, threads, st));  // warm-up
    CUDA_TRY(cudaEventRecord(e0, st));
    CUDA_TRY(launch_smem_gather(table_bytes, iters, pattern, blocks, threads, st));
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e0, e1));
    *lookups = 4.0 * (double)iters * (double)blocks * (double)threads;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HOLDEM_OK;
}

int holdem_microbench_int_issue(int iters, int mix, int blocks_per_sm, int threads, float *elapsed_ms,
                                double *warp_instructions, void *stream) {
    DeviceState *S = nullptr;
    int rc = current_state(&S);
    if (rc) return rc;
    if (iters < 1 || mix < 0 || mix > 2 || threads < 32 || threads > 1024 || (threads & 31) || blocks_per_sm < 1 ||
        !elapsed_ms || !warp_instructions)
        return fail(HOLDEM_ERR_BAD_ARG, "bad micro-benchmark arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = S->sms * blocks_per_sm;
    CUDA_TRY(launch_int_issue(iters, mix, blocks, threads, st));  // warm-up
    CUDA_TRY(cudaEventRecord(e0, st));
    CUDA_TRY(launch_int_issue(iters, mix, blocks, threads, st));
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e0, e1));
    *warp_instructions = 16.0 * (double)iters * (double)blocks * (double)(threads / 32);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HOLDEM_OK;
}

}  // extern "C"
