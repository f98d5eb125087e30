// capi.cu -- C ABI (include/schic_mh.h, include/schic_mh_cuda.h) over the sm_100a kernels.
//
// Host-side orchestration of the hot path: device memory for the fitted set (CSR + signatures),
// chunked H2D upload overlapped with K1 on a second stream, query blocking of the count matrix,
// stage timing with CUDA events, and the D2H of the neighbour lists. There is no CPU compute
// path in this file: without a usable GPU every entry point fails with SCHIC_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/schic_mh_cuda.h"
#include "common.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            cudaGetLastError();                                                                   \
            return fail(e__ == cudaErrorMemoryAllocation ? SCHIC_ERR_OOM : SCHIC_ERR_CUDA,        \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                     \
        }                                                                                         \
    } while (0)

#define ST_TRY(expr)                \
    do {                            \
        int s__ = (expr);           \
        if (s__ != 0) return s__;   \
    } while (0)

// Owning device buffer that can grow while keeping its contents.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    int64_t cap = 0;   // elements
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    int reserve(int64_t n, int64_t keep, cudaStream_t s) {
        if (n <= cap) return 0;
        int64_t ncap = std::max<int64_t>(n, cap + cap / 2);
        T* np = nullptr;
        CU_TRY(cudaMalloc(&np, (size_t)ncap * sizeof(T)));
        if (keep > 0 && p) {
            cudaError_t e = cudaMemcpyAsync(np, p, (size_t)keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cudaFree(np); return fail(SCHIC_ERR_CUDA, cudaGetErrorString(e)); }
        }
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return 0;
    }
};

enum Stage { ST_H2D = 0, ST_SIG, ST_COLL, ST_SELECT, ST_RERANK, ST_GRAPH, ST_D2H, ST_PREP, ST_COUNT };
// NVTX range names of the stages (host-side enqueue ranges; a timeline tool attributes the kernels launched inside)
const char* const kStageName[ST_COUNT] = {"schic:h2d", "schic:K1 signatures", "schic:K3 collisions", "schic:select",
                                          "schic:K4 rerank", "schic:K5 graph", "schic:d2h", "schic:rerank prep"};

}  // namespace

namespace schic {
int set_error(int code, const std::string& msg) { return fail(code, msg); }
// hostconv.cpp
bool narrow_f64_u8(const double* s, int64_t n, uint8_t* d);
bool narrow_f64_u16(const double* s, int64_t n, uint16_t* d);
bool narrow_f32_u8(const float* s, int64_t n, uint8_t* d);
bool narrow_f32_u16(const float* s, int64_t n, uint16_t* d);
void convert_f64_f32(const double* s, int64_t n, float* d);
bool narrow_ids64(const uint64_t* s, int64_t n, uint32_t* lo);
}  // namespace schic

namespace {

// Page-locked host buffer that can grow (contents are not kept).
struct PinBuf {
    unsigned char* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8;
        if (cudaHostAlloc((void**)&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return fail(SCHIC_ERR_OOM, "cudaHostAlloc of the upload staging buffer failed");
        }
        cap = want;
        return 0;
    }
};

// schic_mh_fit_csr_host: the caller's arrays are pageable and (values) float64 -- what scipy holds. A small pool of
// worker threads stages them chunk by chunk into page-locked buffers in wire format (ids: 32-bit words; values: uint8 /
// uint16 counts when every value of the chunk is such an integer, else float32) while the calling thread uploads the
// chunks that are ready: staging, PCIe and K1 overlap. Work items are taken in upload order: all id chunks, then all
// value chunks.
struct Stager {
    static constexpr int64_t kChunk = (int64_t)1 << 20;      // non-zeros per work item
    const void* indices = nullptr; int index_bits = 32;
    const void* data = nullptr; int data_kind = -1;          // -1 none, 0 f32, 1 u16, 2 u8, 3 f64
    int64_t first = 0, nnz = 0, n_chunks = 0;
    uint32_t* s_ids = nullptr;                               // [nnz]
    unsigned char* s_val = nullptr;                          // 4 bytes per non-zero: chunk c starts at byte 4 * c * kChunk
    std::vector<std::atomic<int>> ids_done, val_done;
    std::vector<int8_t> val_kind;                            // wire kind of every value chunk: 0 f32, 1 u16, 2 u8
    std::atomic<int64_t> next{0};
    std::atomic<int> level{2};                               // narrowest kind still worth trying (2 u8, 1 u16, 0 f32)
    std::atomic<bool> any_hi{false};
    std::vector<std::thread> threads;

    int64_t chunk_begin(int64_t c) const { return c * kChunk; }
    int64_t chunk_end(int64_t c) const { return std::min(nnz, (c + 1) * kChunk); }

    void stage_ids(int64_t c) {
        const int64_t a = chunk_begin(c), b = chunk_end(c);
        if (index_bits == 32) {
            memcpy(s_ids + a, (const uint32_t*)indices + first + a, (size_t)(b - a) * 4);
        } else if (schic::narrow_ids64((const uint64_t*)indices + first + a, b - a, s_ids + a)) {
            any_hi.store(true);
        }
        ids_done[c].store(1, std::memory_order_release);
    }
    void stage_val(int64_t c) {
        const int64_t a = chunk_begin(c), b = chunk_end(c), n = b - a;
        unsigned char* dst = s_val + 4 * (size_t)a;
        int kind = 0;
        if (data_kind == 1 || data_kind == 2) {              // already narrow: copy
            kind = data_kind;
            const size_t w = data_kind == 1 ? 2 : 1;
            memcpy(dst, (const unsigned char*)data + (size_t)(first + a) * w, (size_t)n * w);
        } else {
            const double* d64 = data_kind == 3 ? (const double*)data + first + a : nullptr;
            const float* d32 = data_kind == 0 ? (const float*)data + first + a : nullptr;
            int lv = level.load();
            bool ok = false;
            if (lv >= 2) {
                ok = d64 ? schic::narrow_f64_u8(d64, n, dst) : schic::narrow_f32_u8(d32, n, dst);
                if (ok) kind = 2; else { int e = 2; level.compare_exchange_strong(e, 1); lv = 1; }
            }
            if (!ok && lv >= 1) {
                ok = d64 ? schic::narrow_f64_u16(d64, n, (uint16_t*)dst) : schic::narrow_f32_u16(d32, n, (uint16_t*)dst);
                if (ok) kind = 1; else { int e = 1; level.compare_exchange_strong(e, 0); }
            }
            if (!ok) {
                kind = 0;
                if (d64) schic::convert_f64_f32(d64, n, (float*)dst); else memcpy(dst, d32, (size_t)n * 4);
            }
        }
        val_kind[c] = (int8_t)kind;
        val_done[c].store(1, std::memory_order_release);
    }
    void worker() {
        const int64_t items = n_chunks * (data ? 2 : 1);
        for (;;) {
            const int64_t it = next.fetch_add(1);
            if (it >= items) return;
            if (it < n_chunks) stage_ids(it); else stage_val(it - n_chunks);
        }
    }
    void start(int n_threads) {
        n_chunks = (nnz + kChunk - 1) / kChunk;
        ids_done = std::vector<std::atomic<int>>(n_chunks);
        val_done = std::vector<std::atomic<int>>(n_chunks);
        for (auto& x : ids_done) x.store(0);
        for (auto& x : val_done) x.store(0);
        val_kind.assign(n_chunks, 0);
        n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_chunks));
        for (int t = 0; t < n_threads; ++t) threads.emplace_back([this] { worker(); });
    }
    static void wait(std::atomic<int>& f) { while (!f.load(std::memory_order_acquire)) std::this_thread::yield(); }
    void wait_ids_upto(int64_t pos) {                        // ids of positions [0, pos) are staged
        const int64_t last = pos <= 0 ? -1 : (pos - 1) / kChunk;
        for (int64_t c = 0; c <= last; ++c) wait(ids_done[c]);
    }
    void join() { for (auto& t : threads) if (t.joinable()) t.join(); threads.clear(); }
    ~Stager() { join(); }
};

}  // namespace

struct schic_mh_handle {
    schic_mh_params p{};
    int device = 0;
    cudaStream_t stream = nullptr;       // compute stream
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D uploads overlapped with K1
    int64_t launches = 0;

    // local fitted rows
    int64_t n = 0, nnz = 0;
    bool borrowed = false;
    const int64_t* b_indptr = nullptr; const uint32_t* b_feat = nullptr; const float* b_val = nullptr;
    DevBuf<int64_t> indptr; DevBuf<uint32_t> feat; DevBuf<uint32_t> feat_hi; DevBuf<float> val;
    bool have_val = false, have_hi = false;
    DevBuf<uint32_t> sig;
    // SURVEY 8(f)-4 (hyperopt sweeps over -nh): h_j depends only on j, so a prefix of the fitted signatures IS the
    // signature matrix of a smaller hash count. While a prefix is active, sig holds the compacted [n, H_active] copy,
    // sig_fit the fitted [n, H_fit] matrix, and p.number_of_hash_functions == H_active.
    DevBuf<uint32_t> sig_fit; int H_fit = 0; bool sliced = false;
    // block combine ("shingle", SURVEY 8(c) rule 4; off unless params.shingle != 0): K1 fills the raw [n, H] matrix, the
    // combine kernel folds `block` minima per table into sig [n, H / block] and the handle then looks like one created
    // with H / block hash functions -- the same two-buffer state as an active-hash prefix (sliced), undone by the next fit.
    int block = 1;
    DevBuf<unsigned char> k1_scratch;    // work-item list of the filtered / table K1 kernels
    // shared-table K1 (k1_table.cu): per-id presence bytes and slots, hit entries, {entries used, overflow, max id}
    DevBuf<unsigned char> k1t_present; DevBuf<unsigned long long> k1t_slot; DevBuf<uint16_t> k1t_entries;
    DevBuf<uint32_t> k1t_counters;
    bool k1t_prebuilt = false;                // the table for this batch was built ahead of the upload (phase 1)
    // A table built for EVERY id of the hinted range (phase 1) does not depend on the data, only on (range, H, hash width)
    // and -- for speed, not for exactness -- on the mean row length behind its threshold: it is kept and re-used by later
    // fits of the same shape (repeated builds, partial_fit chunks, every rank of a sharded run). `seen_*`: the key of the
    // last batch that had to build from its own ids; the second batch of that shape builds the whole-range table.
    struct { bool valid = false; uint32_t max_id = 0; int H = 0, width = 0; double mean_row = 0.0;
             uint32_t seen_max_id = 0; int seen_H = 0, seen_width = 0; } k1t_cache;
    int k1_table_state = 0;                   // K1 of the last fit: 0 no table, 1 built from the batch's ids, 2 whole-range table built, 3 cached table re-used
    std::vector<uint32_t> host_lo, host_hi;   // split 64-bit ids of the last fit (must outlive an async upload)
    DevBuf<unsigned char> val_stage;          // narrow (uint16 / uint8) counts as uploaded, before widening
    PinBuf pin_ids, pin_val;                  // page-locked staging of schic_mh_fit_csr_host (ids, values in wire format)
    int64_t count_budget = 0;                 // elements of the uint16 count matrix we allow ourselves (cached)
    uint64_t feature_range = 0;               // caller's hint: all ids < feature_range (0: unknown, read the maximum back)
    DevBuf<double> norm2; int64_t norm_rows = 0; const void* norm_of = nullptr;
    DevBuf<uint32_t> win_dir, max_feat; int n_windows = 0;   // window directory (2^16-id windows) of the rerank (0: hash kernel)
    int window_mode = 32;
    // counts rerank (k4c_rerank.cu): packed stream (id mod 2^16) << 16 | count of the prepared database, the prep flags
    // {not packable, #values >= 255, id beyond the range, -} and the uint64 dot-product accumulator
    DevBuf<uint32_t> packed, k4_flags; DevBuf<unsigned long long> dotacc;
    bool packed_ok = false; int64_t prep_nnz = 0;
    int64_t b_end = 0;                        // end position of a borrowed CSR (indptr[n])
    int fit_data_kind = 0;                    // how the values of the last fit arrived: 0 float32, 1 uint16, 2 uint8
    int rerank_path = 0;                      // what the last rerank launched: 1 counts only, 2 counts + generic behind it, 3 generic only

    // external database (distributed query)
    bool ext = false;
    int64_t db_n = 0, db_row0 = 0;
    const uint32_t* db_sig = nullptr; const int64_t* db_indptr = nullptr;
    const uint32_t* db_feat = nullptr; const float* db_val = nullptr;
    cudaEvent_t db_ready = nullptr;      // caller's event: the external CSR is complete once it has fired
    const double* ext_norm2 = nullptr; const uint32_t* ext_dir = nullptr; int ext_nw = 0;   // caller-provided rerank prep
    const uint32_t* ext_packed = nullptr; const uint32_t* ext_flags = nullptr; int64_t ext_nnz = 0;

    // scratch of kneighbors
    DevBuf<uint16_t> counts; DevBuf<int32_t> cand_idx, cand_cnt, cand_nv, sel_flags, part_idx; DevBuf<float> part_dist;
    DevBuf<uint32_t> sig_masked;   // database signatures with oversized buckets blanked (K2)
    DevBuf<uint32_t> k3h_table, k3h_sizes; DevBuf<uint16_t> sig16;   // packed K3: renaming tables, bucket sizes, renamed signatures
    const void* sig16_of = nullptr; int64_t sig16_rows = 0;          // database the renamed signatures were made from (block API)
    DevBuf<uint32_t> sig16T;                                         // the same, transposed, for the TMA-staged K3
    alignas(64) unsigned char k3_tmap[128]; bool k3_tma = false;     // tensor map over sig16T (valid for the last rename)
    const uint16_t* given_counts = nullptr; int64_t given_ld = 0;    // caller-assembled count matrix for the next kneighbors call
    bool coll_blocks_open = false;                                   // ST_COLL holds block launches of the build in progress
    DevBuf<unsigned long long> long_keys;                            // rows beyond the shared-memory sorts (k6_longrows.cu)
    DevBuf<uint32_t> feat_rank; DevBuf<unsigned char> rank_scratch;  // ids >= 2^32: rank of every id among the distinct ids
    bool rank_valid = false; uint32_t rank_max = 0;
    DevBuf<int32_t> out_idx, out_nv; DevBuf<float> out_dist;
    DevBuf<int64_t> g_indptr; DevBuf<int32_t> g_indices; DevBuf<float> g_data;

    // stage timing: accumulated event pairs of the last fit / kneighbors
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[ST_COUNT];
    std::vector<cudaEvent_t> pool;      // every event ever created (owned)
    std::vector<cudaEvent_t> free_ev;   // recycled
    std::vector<cudaEvent_t> misc_ev;   // ordering events of the current fit (not stage timers)
    cudaEvent_t val_done = nullptr;     // pending async value upload (member of misc_ev)

    const int64_t* d_indptr() const { return borrowed ? b_indptr : indptr.p; }
    const uint32_t* d_feat() const { return borrowed ? b_feat : feat.p; }
    const float* d_val() const { return borrowed ? b_val : (have_val ? val.p : nullptr); }
};

namespace {

constexpr int RERANK_CHUNK = 1024;   // candidates per K4 launch when a list exceeds one CTA's shared memory


int use_device(const schic_mh_handle* h) {
    CU_TRY(cudaSetDevice(h->device));
    return 0;
}

cudaEvent_t get_event(schic_mh_handle* h) {
    if (!h->free_ev.empty()) {
        cudaEvent_t e = h->free_ev.back();
        h->free_ev.pop_back();
        return e;
    }
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    h->pool.push_back(e);
    return e;
}

cudaEvent_t get_misc_event(schic_mh_handle* h) {
    cudaEvent_t e = get_event(h);
    if (e) h->misc_ev.push_back(e);
    return e;
}

struct StageTimer {
    schic_mh_handle* h; Stage st; cudaStream_t s; cudaEvent_t a;
    StageTimer(schic_mh_handle* hh, Stage stage, cudaStream_t stream) : h(hh), st(stage), s(stream) {
        nvtxRangePushA(kStageName[st]);
        a = get_event(h);
        if (a) cudaEventRecord(a, s);
    }
    ~StageTimer() {
        cudaEvent_t b = get_event(h);
        if (a && b) { cudaEventRecord(b, s); h->ev[st].push_back({a, b}); }
        nvtxRangePop();
    }
};

void reset_stage(schic_mh_handle* h, std::initializer_list<Stage> stages) {
    for (Stage s : stages) {
        for (auto& pr : h->ev[s]) { h->free_ev.push_back(pr.first); h->free_ev.push_back(pr.second); }
        h->ev[s].clear();
    }
}

int check(const schic_mh_handle* h) {
    if (!h) return fail(SCHIC_ERR_INVALID_ARG, "null handle");
    return 0;
}

void add_launches(schic_mh_handle* h, int n) { if (n > 0) h->launches += n; }

// K1 on rows [r0, r1) of the handle's CSR (positions are absolute in the device arrays).
// SCHIC_K1_MODE (experiments, tests): "chunk" = the unfiltered fixed-chunk kernel, "tile" / "direct" = the
// filtered per-cell kernel with / without the shared-memory tile, "table" = force the shared-table path;
// default: shared table when the batch re-uses feature ids enough (k1t_applicable), else filtered "tile".
enum K1Mode { K1_AUTO = 0, K1_CHUNK, K1_TILE, K1_DIRECT, K1_TABLE };
K1Mode k1_mode_from_env() {
    const char* m = getenv("SCHIC_K1_MODE");
    if (!m) return K1_AUTO;
    if (strcmp(m, "chunk") == 0) return K1_CHUNK;
    if (strcmp(m, "tile") == 0) return K1_TILE;
    if (strcmp(m, "direct") == 0) return K1_DIRECT;
    if (strcmp(m, "table") == 0) return K1_TABLE;
    return K1_AUTO;
}

// {entries used, table overflow, max id scratch, id-beyond-range flag (sticky)}
int ensure_counters(schic_mh_handle* h) {
    if (h->k1t_counters.cap >= 4) return 0;
    ST_TRY(h->k1t_counters.reserve(4, 0, h->stream));
    CU_TRY(cudaMemsetAsync(h->k1t_counters.p, 0, 4 * sizeof(uint32_t), h->stream));
    return 0;
}

uint32_t esc_limit_for(int64_t nnz) { return (uint32_t)std::min<int64_t>(nnz / 64, 0x7FFFFFFF); }

// After a host synchronisation: did a kernel see an id beyond the caller's feature-range hint (K1 table path: counters[3];
// rerank preparation: flags[2])? Was a database that came as packed counts only (no ids / values for the generic
// kernels) not packable after all?
int check_range_flag(schic_mh_handle* h) {
    if (h->ext && h->ext_packed && h->ext_flags && !h->db_val && h->rerank_path == 1) {
        uint32_t f[4] = {0, 0, 0, 0};
        CU_TRY(cudaMemcpyAsync(f, h->ext_flags, sizeof(f), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (f[0] != 0 || f[1] > esc_limit_for(h->ext_nnz))
            return fail(SCHIC_ERR_INVALID_ARG, "the database was handed over as packed integer counts only, but its values are not "
                                               "small non-negative integers: exchange the CSR ids and values as well");
        if (f[2]) return fail(SCHIC_ERR_INVALID_ARG, "a feature id is >= the range declared with schic_mh_set_feature_range");
    }
    if (h->feature_range == 0) return 0;
    uint32_t flag = 0, pflag = 0;
    if (h->k1t_counters.cap >= 4)
        CU_TRY(cudaMemcpyAsync(&flag, h->k1t_counters.p + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (h->k4_flags.cap >= 4 && h->norm_of)
        CU_TRY(cudaMemcpyAsync(&pflag, h->k4_flags.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (flag || pflag) {
        if (flag) CU_TRY(cudaMemsetAsync(h->k1t_counters.p + 3, 0, sizeof(uint32_t), h->stream));
        if (pflag) { CU_TRY(cudaMemsetAsync(h->k4_flags.p + 2, 0, sizeof(uint32_t), h->stream)); h->norm_of = nullptr; }
        return fail(SCHIC_ERR_INVALID_ARG, "a feature id is >= the range declared with schic_mh_set_feature_range");
    }
    return 0;
}

bool k1t_cache_hit(const schic_mh_handle* h, uint32_t max_id, int64_t n_rows, int64_t nnz) {
    const auto& c = h->k1t_cache;
    if (!c.valid || c.max_id != max_id || c.H != h->p.number_of_hash_functions || c.width != h->p.hash_width || n_rows <= 0) return false;
    if (getenv("SCHIC_K1_NO_CACHE")) return false;
    const double r = ((double)nnz / (double)n_rows) / c.mean_row;
    return r > 0.6 && r < 1.7;      // the cached threshold still fits: a lower one only sends more (cell, j) to the exact fallback
}
void k1t_cache_store(schic_mh_handle* h, uint32_t max_id, int64_t n_rows, int64_t nnz) {
    h->k1t_cache.valid = true; h->k1t_cache.max_id = max_id; h->k1t_cache.H = h->p.number_of_hash_functions;
    h->k1t_cache.width = h->p.hash_width; h->k1t_cache.mean_row = (double)nnz / (double)std::max<int64_t>(1, n_rows);
}

// cheap test before anything touches the device: could this batch qualify for the shared-table path at all?
bool k1_table_candidate(const schic_mh_handle* h, int64_t n_rows, int64_t nnz) {
    if (h->p.hash_width == 64 && h->have_hi) return false;      // ids >= 2^32: far beyond the per-id arrays
    const K1Mode m = k1_mode_from_env();
    if (m == K1_TABLE) return n_rows > 0 && nnz > 0;
    if (m != K1_AUTO) return false;
    const int H = h->p.number_of_hash_functions;
    return n_rows > 0 && nnz >= ((int64_t)1 << 21) && schic::k1t_tau(n_rows, nnz, H) >= 4.0;
}

// K1 on rows [r0, r1) of the handle's CSR (positions are absolute in the device arrays).
int hash_rows(schic_mh_handle* h, int64_t r0, int64_t r1, int64_t pos0, int64_t pos1, cudaStream_t s) {
    const int H = h->p.number_of_hash_functions;
    uint32_t* sig = h->sig.p + r0 * (int64_t)H;
    const int64_t n_rows = r1 - r0, nnz = pos1 - pos0;
    const K1Mode mode = k1_mode_from_env();
    int nl = -1;
    if (mode == K1_CHUNK && h->p.hash_width == 32) {
        nl = schic::launch_signatures32(h->d_indptr() + r0, h->d_feat(), n_rows, nnz, pos0, H, sig, s);
    } else if (h->k1t_prebuilt) {
        // the batch's table (every id of the hinted range) is already built: any row range of the batch only needs
        // lookup + fallback, so the chunks are hashed as they land
        const uint32_t max_id = (uint32_t)(h->feature_range - 1);
        const size_t need = schic::k1t_items_bytes(n_rows, nnz);
        if ((int64_t)need > h->k1_scratch.cap) {
            CU_TRY(cudaStreamSynchronize(s));      // an earlier chunk's kernels may still read the old scratch
            ST_TRY(h->k1_scratch.reserve((int64_t)need, 0, s));
        }
        nl = schic::launch_signatures_table(h->p.hash_width, h->d_indptr() + r0, h->d_feat(), n_rows, nnz, pos0, max_id, H, sig,
                                            h->k1t_present.p, reinterpret_cast<uint2*>(h->k1t_slot.p), h->k1t_entries.p,
                                            (size_t)h->k1t_entries.cap, h->k1t_counters.p, h->k1_scratch.p, 2, s);
    } else {
        if (k1_table_candidate(h, n_rows, nnz)) {
            // the id range decides: from the caller's hint, else the largest id of the batch is read back
            // (4 bytes, one stream synchronisation per batch)
            ST_TRY(ensure_counters(h));
            uint32_t max_id = 0;
            if (h->feature_range > 0 && h->feature_range <= 0xFFFFFFFFull) {
                max_id = (uint32_t)(h->feature_range - 1);
            } else {
                add_launches(h, schic::launch_max_id(h->d_feat() + pos0, nnz, h->k1t_counters.p + 2, s));
                CU_TRY(cudaMemcpyAsync(&max_id, h->k1t_counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
                CU_TRY(cudaStreamSynchronize(s));
            }
            if (mode == K1_TABLE ? (max_id < 134217728u && H <= 65535) : schic::k1t_applicable(n_rows, nnz, max_id, H)) {
                const size_t ecap = schic::k1t_entries_capacity(n_rows, nnz, max_id, H);
                const bool hinted = h->feature_range > 0 && h->feature_range <= 134217728ull && mode == K1_AUTO;
                auto& c = h->k1t_cache;
                int phase = 0;
                if (hinted && k1t_cache_hit(h, max_id, n_rows, nnz)) {
                    phase = 2;                            // whole-range table of an earlier fit: lookup + fallback only
                    h->k1_table_state = 3;
                } else if (hinted && !getenv("SCHIC_K1_NO_CACHE") &&
                           ((double)h->feature_range * (double)H < 175.0 * (double)nnz ||
                            (c.seen_max_id == max_id && c.seen_H == H && c.seen_width == h->p.hash_width))) {
                    phase = 1;                            // build it for every id of the range, keep it
                    h->k1_table_state = 2;
                } else {
                    c.valid = false;                      // the per-id arrays are about to hold this batch's own table
                    c.seen_max_id = max_id; c.seen_H = H; c.seen_width = h->p.hash_width;
                    h->k1_table_state = 1;
                }
                if (phase != 2) {
                    ST_TRY(h->k1t_present.reserve((int64_t)max_id + 1, 0, s));
                    ST_TRY(h->k1t_slot.reserve((int64_t)max_id + 1, 0, s));
                    ST_TRY(h->k1t_entries.reserve((int64_t)ecap, 0, s));
                }
                ST_TRY(h->k1_scratch.reserve((int64_t)schic::k1t_items_bytes(n_rows, nnz), 0, s));
                nl = 0;
                if (phase == 1) {
                    c.valid = false;
                    const int b = schic::launch_signatures_table(h->p.hash_width, nullptr, nullptr, n_rows, nnz, 0, max_id, H, nullptr,
                                                                 h->k1t_present.p, reinterpret_cast<uint2*>(h->k1t_slot.p),
                                                                 h->k1t_entries.p, ecap, h->k1t_counters.p, nullptr, 1, s);
                    if (b < 0) return fail(SCHIC_ERR_CUDA, "K1 table build failed");
                    nl += b;
                    k1t_cache_store(h, max_id, n_rows, nnz);
                    phase = 2;
                }
                const int r = schic::launch_signatures_table(h->p.hash_width, h->d_indptr() + r0, h->d_feat(), n_rows, nnz, pos0, max_id, H, sig,
                                                             h->k1t_present.p, reinterpret_cast<uint2*>(h->k1t_slot.p),
                                                             h->k1t_entries.p, phase == 2 ? (size_t)h->k1t_entries.cap : ecap,
                                                             h->k1t_counters.p, h->k1_scratch.p, phase, s);
                nl = r < 0 ? -1 : nl + r;
            }
        }
        if (nl < 0 && h->p.hash_width == 64) {
            nl = schic::launch_signatures64(h->d_indptr() + r0, h->d_feat(), h->have_hi ? h->feat_hi.p : nullptr,
                                            n_rows, nnz, pos0, H, sig, s);
        } else if (nl < 0) {
            // the scratch of an earlier chunk may still be read by its kernel: same stream, so reuse is ordered,
            // but growing would free it -- synchronise before that
            const size_t need = schic::k1f_scratch_bytes(n_rows, nnz);
            if ((int64_t)need > h->k1_scratch.cap) {
                CU_TRY(cudaStreamSynchronize(s));
                ST_TRY(h->k1_scratch.reserve((int64_t)need, 0, s));
            }
            nl = schic::launch_signatures32_filter(h->d_indptr() + r0, h->d_feat(), n_rows, nnz, H, sig, h->k1_scratch.p,
                                                   mode == K1_DIRECT ? 1 : 0, s);
        }
    }
    if (nl < 0) return fail(SCHIC_ERR_CUDA, "K1 launch failed");
    add_launches(h, nl);
    CU_TRY(cudaGetLastError());
    return 0;
}

int apply_block_combine(schic_mh_handle* h);
// Undo schic_mh_set_active_hashes: the fitted [n, H_fit] matrix becomes the live one again.
void restore_fitted_signatures(schic_mh_handle* h) {
    if (!h->sliced) return;
    std::swap(h->sig.p, h->sig_fit.p);
    std::swap(h->sig.cap, h->sig_fit.cap);
    h->p.number_of_hash_functions = h->H_fit;
    h->sliced = false;
}

// End of every fit when block combine is on: fold the raw minima (K2) and switch the handle to the H / block tables.
int apply_block_combine(schic_mh_handle* h) {
    if (h->block <= 1 || h->sliced) return SCHIC_OK;
    const int H_raw = h->p.number_of_hash_functions, T = H_raw / h->block;
    cudaStream_t s = h->stream;
    ST_TRY(h->sig_fit.reserve(std::max<int64_t>(1, h->n * (int64_t)T), 0, s));
    {
        StageTimer t(h, ST_SIG, s);
        add_launches(h, schic::launch_block_combine(h->p.hash_width, h->sig.p, h->n, H_raw, h->block, h->sig_fit.p, s));
    }
    CU_TRY(cudaGetLastError());
    h->H_fit = H_raw;
    std::swap(h->sig.p, h->sig_fit.p);      // sig: the keys [n, T]; sig_fit: the raw matrix (restored by the next fit)
    std::swap(h->sig.cap, h->sig_fit.cap);
    h->sliced = true;
    h->p.number_of_hash_functions = T;
    return SCHIC_OK;
}

int prune_buckets(schic_mh_handle* h, const uint32_t* db_sig, int64_t ndb) {
    const int H = h->p.number_of_hash_functions;
    ST_TRY(h->sig_masked.reserve(ndb * (int64_t)H, 0, h->stream));
    add_launches(h, schic::launch_bucket_mask(db_sig, ndb, H, h->p.max_bin_size, h->sig_masked.p, h->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

// Feature ids beyond 32 bits: the rerank runs on the rank of every id among the distinct ids of the fitted rows (a
// monotone relabelling, so equality and the order inside a row are preserved). Once per fit; one read-back (the number
// of distinct ids sizes the window directory).
int ensure_rank_labels(schic_mh_handle* h) {
    if (h->rank_valid) return 0;
    if (h->nnz > INT32_MAX) return fail(SCHIC_ERR_UNSUPPORTED, "exact rerank with ids >= 2^32 supports up to 2^31 non-zeros");
    cudaStream_t s = h->stream;
    CU_TRY(cudaStreamSynchronize(h->copy_stream));
    const size_t need = schic::relabel_scratch_bytes(h->nnz);
    ST_TRY(h->feat_rank.reserve(h->nnz + 4, 0, s));
    ST_TRY(h->rank_scratch.reserve((int64_t)need, 0, s));
    ST_TRY(h->max_feat.reserve(1, 0, s));
    const int nl = schic::launch_relabel_ids(h->feat.p, h->feat_hi.p, h->nnz, h->feat_rank.p, h->max_feat.p, h->rank_scratch.p,
                                             need, s);
    if (nl < 0) return fail(SCHIC_ERR_CUDA, "id relabelling failed");
    add_launches(h, nl);
    CU_TRY(cudaMemcpyAsync(&h->rank_max, h->max_feat.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    h->rank_scratch.release();
    h->rank_valid = true;
    return 0;
}

// Per-database preparation of the exact rerank, cached until the next fit -- ONE pass over ids + values
// (k4c_prep_kernel): fp64 row norms and, when the id range spans at most kMaxWindows windows of 2^16 ids, the window
// directory, the packed stream of the counts kernel and the flags that say whether the data is packable.
// SCHIC_K4_MODE: "hash" = norms only (hash-table kernel), "window" / "window64" = directory, generic windowed kernel
// only, default = counts kernel with the generic kernels behind it (tests compare them all).
int ensure_norms(schic_mh_handle* h, const int64_t* indptr, const uint32_t* feat, const float* val, int64_t rows, int64_t nnz_end) {
    if (h->norm_of == (const void*)indptr && h->norm_rows == rows) return 0;
    constexpr int kMaxWindows = 8192;
    cudaStream_t s = h->stream;
    if (h->val_done) CU_TRY(cudaStreamWaitEvent(s, h->val_done, 0));   // async fit: the values may still be in flight
    ST_TRY(h->norm2.reserve(rows, 0, s));
    ST_TRY(h->k4_flags.reserve(4, 0, s));
    CU_TRY(cudaMemsetAsync(h->k4_flags.p, 0, 4 * sizeof(uint32_t), s));
    h->n_windows = 0;
    h->packed_ok = false;
    h->prep_nnz = 0;
    const char* mode = getenv("SCHIC_K4_MODE");
    h->window_mode = (mode && strcmp(mode, "window64") == 0) ? 64 : 32;
    int nw = 0;
    if (!(mode && strcmp(mode, "hash") == 0)) {
        uint32_t mx = 0;
        if (h->have_hi && !h->ext && h->rank_valid) {
            mx = h->rank_max;                      // ids were replaced by their ranks
        } else if (h->feature_range > 0 && h->feature_range <= 0xFFFFFFFFull) {
            mx = (uint32_t)(h->feature_range - 1);
        } else {
            ST_TRY(h->max_feat.reserve(1, 0, s));
            const int nl = schic::launch_row_max_feature(indptr, feat, rows, h->max_feat.p, s);
            if (nl < 0) return fail(SCHIC_ERR_CUDA, "max-feature launch failed");
            add_launches(h, nl);
            CU_TRY(cudaMemcpyAsync(&mx, h->max_feat.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
        }
        nw = schic::k4c_window_count(mx);
        if (nw > kMaxWindows) nw = 0;
    }
    const bool want_packed = nw > 0 && !(mode && (strcmp(mode, "window") == 0 || strcmp(mode, "window64") == 0));
    if (want_packed && nnz_end < 0) {      // external database without its own preparation: one read-back of indptr[rows]
        CU_TRY(cudaMemcpyAsync(&nnz_end, indptr + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
    }
    if (nw > 0) ST_TRY(h->win_dir.reserve(rows * (int64_t)(nw + 1), 0, s));
    if (want_packed) ST_TRY(h->packed.reserve(nnz_end + 4, 0, s));      // slices are read as aligned 16-byte words
    add_launches(h, schic::launch_rerank_prep(indptr, feat, val, rows, nw, h->norm2.p, nw > 0 ? h->win_dir.p : nullptr,
                                              want_packed ? h->packed.p : nullptr, h->k4_flags.p, s));
    CU_TRY(cudaGetLastError());
    h->n_windows = nw;
    h->packed_ok = want_packed;
    h->prep_nnz = want_packed ? nnz_end : 0;
    h->norm_of = indptr;
    h->norm_rows = rows;
    return 0;
}

// Packed K3 (two hash functions per compare): rename the database signatures to 16-bit ids. Returns 1 when the
// packed kernels can be used (h->sig16 filled), 0 when not applicable; SCHIC_K3_MODE=plain forces 0.
bool packed_counts_possible(const schic_mh_handle* h, int64_t ndb) {
    const char* mode = getenv("SCHIC_K3_MODE");
    return !(mode && strcmp(mode, "plain") == 0) && schic::k3h_applicable(ndb, h->p.number_of_hash_functions);
}

// Packed K3 on database rows [q_row0, +nq) x [d_row0, +nd) of the renamed signatures (TMA-staged when available)
int launch_k3_packed(schic_mh_handle* h, int64_t q_row0, int64_t nq, int64_t d_row0, int64_t nd, uint16_t* cnt, int64_t ld,
                     uint16_t* cnt_t, int64_t ld_t) {
    const int H = h->p.number_of_hash_functions;
    // (a TMA box must start on a 16-byte boundary of the row axis: row offsets that are not multiples of 4 take the LDG kernel)
    if (h->k3_tma && (q_row0 & 3) == 0 && (d_row0 & 3) == 0)
        return schic::launch_collision_counts_tma(h->k3_tmap, q_row0, nq, d_row0, nd, H, cnt, ld, cnt_t, ld_t, h->stream);
    const int64_t Hp = schic::k3h_padded_hashes(H);
    const uint16_t* q16 = h->sig16.p + q_row0 * Hp;
    const uint16_t* d16 = h->sig16.p + d_row0 * Hp;
    if (cnt_t && !(q_row0 == d_row0 && nq == nd))
        return schic::launch_collision_counts_packed_t(q16, nq, d16, nd, H, cnt, ld, cnt_t, ld_t, h->stream);
    return schic::launch_collision_counts_packed(q16, nq, d16, nd, H, cnt, ld, h->stream);
}

// db_sig: the UNPRUNED database signatures; buckets with >= max_bin_size members are blanked here when ndb reaches it
int prepare_packed_counts(schic_mh_handle* h, const uint32_t* db_sig, int64_t ndb) {
    const int H = h->p.number_of_hash_functions;
    if (!packed_counts_possible(h, ndb)) return 0;
    const bool prune = ndb >= h->p.max_bin_size;
    ST_TRY(h->k3h_table.reserve((int64_t)schic::k3h_table_elems(ndb, H), 0, h->stream));
    if (prune) ST_TRY(h->k3h_sizes.reserve((int64_t)schic::k3h_table_elems(ndb, H), 0, h->stream));
    ST_TRY(h->sig16.reserve(ndb * (int64_t)schic::k3h_padded_hashes(H), 0, h->stream));
    const int nl = schic::launch_rename_signatures(db_sig, ndb, H, h->k3h_table.p, h->sig16.p,
                                                   prune ? h->k3h_sizes.p : nullptr, prune ? h->p.max_bin_size : 0, h->stream);
    if (nl < 0) return fail(SCHIC_ERR_CUDA, "signature renaming launch failed");
    add_launches(h, nl);
    h->sig16_of = nullptr;          // (the block API re-validates what it needs)
    // TMA-staged K3 (default; SCHIC_K3_TMA=0: LDG-staged kernel): transposed copy + tensor map
    h->k3_tma = false;
    const char* tma = getenv("SCHIC_K3_TMA");
    if (!(tma && atoi(tma) == 0) && ndb < INT32_MAX) {
        ST_TRY(h->sig16T.reserve((int64_t)schic::k3t_transposed_elems(ndb, H), 0, h->stream));
        if (schic::k3t_make_tensor_map(h->sig16T.p, ndb, H, h->k3_tmap) == 0) {
            add_launches(h, schic::launch_transpose_signatures(h->sig16.p, ndb, H, h->sig16T.p, h->stream));
            h->k3_tma = true;
        }
    }
    return 1;
}

// SCHIC_FORCE_LONGROWS=1 (tests): take the global-memory row sorts even when a row fits in shared memory
bool force_longrows() { const char* e = getenv("SCHIC_FORCE_LONGROWS"); return e && atoi(e) != 0; }

// scratch of the global-memory row sorts: rows per pass for rows of `row_len` keys (<= 1 GiB of keys at a time)
int64_t long_rows_per_pass(int64_t row_len, int64_t nq) {
    const int64_t P = schic::longrow_pow2(row_len);
    return std::max<int64_t>(1, std::min<int64_t>(nq, ((int64_t)1 << 27) / P));
}
int reserve_long_keys(schic_mh_handle* h, int64_t row_len, int64_t nq, int64_t* rows_per_pass) {
    *rows_per_pass = long_rows_per_pass(row_len, nq);
    return h->long_keys.reserve(*rows_per_pass * (int64_t)schic::longrow_pow2(row_len), 0, h->stream);
}

// K4 over candidate lists [nq, kcand] (cand_idx == nullptr: brute force, every database row is a candidate of every
// query). Lists that fit one CTA's shared memory take one launch; longer ones (and brute force) are reranked in chunks
// of RERANK_CHUNK candidates - each chunk leaves its best min(k, chunk) sorted in a scratch - and merged per row.
// Every launch is the counts kernel (when the database has a packed stream) with the generic float kernels behind it:
// the prep flags on the device decide which of the two does the work, the other exits at once.
int rerank_lists(schic_mh_handle* h, const int64_t* db_indptr, const uint32_t* db_feat, const float* db_val, int64_t ndb,
                 bool ext_prep, int64_t row0, int64_t nq, const int32_t* cand_idx, const int32_t* cand_nv, int kcand, int k,
                 bool exclude_self, int32_t* d_idx, float* d_dist, int32_t* d_nv) {
    cudaStream_t s = h->stream;
    StageTimer t(h, ST_RERANK, s);
    const double* r_norm2 = ext_prep ? h->ext_norm2 : h->norm2.p;
    const uint32_t* r_dir = ext_prep ? h->ext_dir : (h->n_windows > 0 ? h->win_dir.p : nullptr);
    const int r_nw = ext_prep ? h->ext_nw : h->n_windows, r_mode = ext_prep ? 32 : h->window_mode;
    const uint32_t* r_packed = ext_prep ? h->ext_packed : (h->packed_ok ? h->packed.p : nullptr);
    const uint32_t* r_flags = ext_prep ? h->ext_flags : (h->packed_ok ? h->k4_flags.p : nullptr);
    const int64_t r_nnz = ext_prep ? h->ext_nnz : h->prep_nnz;
    const char* mode = getenv("SCHIC_K4_MODE");
    const bool forced_generic = mode && (strcmp(mode, "hash") == 0 || strcmp(mode, "window") == 0 || strcmp(mode, "window64") == 0);
    const bool counts = r_packed && r_flags && r_dir && r_nw > 0 && !forced_generic;
    const bool have_generic = db_feat && db_val;
    // uint8 counts straight from the host are packable by construction: nothing to fall back to
    const bool generic = have_generic && !(counts && !h->ext && !h->borrowed && h->fit_data_kind == 2);
    if (!counts && !have_generic) return fail(SCHIC_ERR_INVALID_ARG, "exact rerank needs the CSR ids and values (or a packed database)");
    const uint32_t esc_limit = esc_limit_for(r_nnz);
    int64_t group_bytes = (int64_t)80 << 20;         // share of the packed stream one window group spans (L2-resident; measured: 20/40/80/160 MB -> 17.9/16.8/16.2/16.5 ms)
    if (const char* g = getenv("SCHIC_K4_GROUP_MB")) group_bytes = (int64_t)atoll(g) << 20;              // 0: one group
    if (const char* g = getenv("SCHIC_K4_GROUP_WINDOWS")) group_bytes = -std::max<int64_t>(1, atoll(g));  // windows per group
    h->rerank_path = counts ? (generic ? 2 : 1) : 3;
    if (counts) ST_TRY(h->dotacc.reserve((int64_t)schic::k4c_dotacc_elems(nq, std::min(kcand, RERANK_CHUNK)), 0, s));

    // one launch (counts and / or generic) over candidate columns [c0, c0 + cn); -1: does not fit, chunk the lists
    auto launch_one = [&](int c0, int cn, int kout, int32_t* o_idx, float* o_dist, int o_stride, int o_off, int final,
                          int n_all, int allow_hash) -> int {
        int nl = 0;
        bool did_counts = false;
        if (counts) {
            const int r = schic::launch_rerank_counts(db_indptr, r_packed, r_norm2, r_dir, r_nw, r_nnz, group_bytes, row0, nq,
                                                      cand_idx, cand_nv, kcand, c0, cn, kout, o_idx, o_dist, o_stride, o_off,
                                                      final, d_nv, n_all, r_flags, esc_limit, h->dotacc.p, s);
            if (r < 0 && !generic) return -1;
            if (r >= 0) { nl += r; did_counts = true; }
        }
        if (generic) {
            const int r = schic::launch_rerank(db_indptr, db_feat, db_val, 0, r_norm2, r_dir, r_nw, r_mode, row0, nq, cand_idx,
                                               cand_nv, kcand, c0, cn, kout, o_idx, o_dist, o_stride, o_off, final, d_nv, n_all,
                                               allow_hash, did_counts ? r_flags : nullptr, esc_limit, s);
            if (r < 0) return -1;
            nl += r;
        }
        return nl;
    };

    int nl = -1;
    if (cand_idx && (!counts || kcand <= RERANK_CHUNK))
        nl = launch_one(0, kcand, k, d_idx, d_dist, k, 0, 1, 0,
                        /* a long list: the windowed kernel on chunks beats the hash-table kernel on the whole */
                        (r_dir && r_nw > 0 && kcand > RERANK_CHUNK) ? 0 : 1);
    if (nl < 0) {
        const int kc = (int)std::min<int64_t>((int64_t)k + (exclude_self ? 1 : 0), RERANK_CHUNK);
        const int nchunks = (kcand + RERANK_CHUNK - 1) / RERANK_CHUNK;
        const int64_t stride = (int64_t)nchunks * kc;
        if (stride > INT32_MAX / 2) return fail(SCHIC_ERR_UNSUPPORTED, "rerank: candidate lists too long");
        ST_TRY(h->part_idx.reserve(nq * stride, 0, s));
        ST_TRY(h->part_dist.reserve(nq * stride, 0, s));
        nl = 0;
        for (int ci = 0; ci < nchunks; ++ci) {
            const int c0 = ci * RERANK_CHUNK, cn = std::min(RERANK_CHUNK, kcand - c0);
            const int r = launch_one(c0, cn, kc, h->part_idx.p, h->part_dist.p, (int)stride, ci * kc, 0, (int)ndb, 1);
            if (r < 0) return fail(SCHIC_ERR_UNSUPPORTED, "rerank: chunk does not fit in shared memory");
            nl += r;
        }
        int r = (stride > 16384 || force_longrows()) ? -1
                               : schic::launch_rerank_merge(h->part_idx.p, h->part_dist.p, (int)stride, cand_idx ? cand_nv : nullptr,
                                                            kcand, (int)ndb, exclude_self ? row0 : -1, nq, k, d_idx, d_dist, d_nv, s);
        if (r < 0) {           // more partial results per row than one CTA's shared memory sorts: global sort
            int64_t per = 0;
            ST_TRY(reserve_long_keys(h, stride, nq, &per));
            r = schic::launch_rerank_merge_long(h->part_idx.p, h->part_dist.p, (int)stride, cand_idx ? cand_nv : nullptr, kcand,
                                                (int)ndb, exclude_self ? row0 : -1, nq, k, d_idx, d_dist, d_nv, h->long_keys.p,
                                                per, s);
        }
        nl += r;
    }
    add_launches(h, nl);
    CU_TRY(cudaGetLastError());
    return 0;
}

// The whole query side on device. Results land in h->out_idx / out_dist / out_nv ([n, k]) unless
// caller-provided device outputs are given.
int kneighbors_on_device(schic_mh_handle* h, int k, int fast_arg, int32_t* d_idx, float* d_dist, int32_t* d_nv) {
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (k < 1) return fail(SCHIC_ERR_INVALID_ARG, "k must be >= 1");
    const bool fast = fast_arg < 0 ? h->p.fast != 0 : fast_arg != 0;
    const int H = h->p.number_of_hash_functions;
    const int64_t nq = h->n;
    const int64_t ndb = h->ext ? h->db_n : h->n;
    const int64_t row0 = h->ext ? h->db_row0 : 0;
    const uint32_t* db_sig = h->ext ? h->db_sig : h->sig.p;
    const int64_t* db_indptr = h->ext ? h->db_indptr : h->d_indptr();
    const uint32_t* db_feat_in = h->ext ? h->db_feat : h->d_feat();
    const uint32_t* db_feat = db_feat_in;      // (replaced by the ids' ranks when ids need more than 32 bits)
    const float* db_val = h->ext ? h->db_val : h->d_val();
    if (k > ndb) k = (int)ndb;   // cannot have more neighbours than fitted rows; outputs keep stride k
    const int64_t kc64 = fast ? (int64_t)k : std::min<int64_t>((int64_t)k * std::max(1, h->p.excess_factor), ndb);
    const int kcand = (int)kc64;
    const bool packed_db = h->ext && h->ext_packed && h->ext_flags && h->ext_norm2 && h->ext_dir && h->ext_nw > 0;
    if (!fast) {
        if ((!db_val || !db_feat) && !packed_db) return fail(SCHIC_ERR_INVALID_ARG, "exact rerank needs the CSR values (fit with data)");
        if (h->have_hi) {
            if (h->ext || h->borrowed) return fail(SCHIC_ERR_UNSUPPORTED, "exact rerank with ids >= 2^32 needs the rows fitted through the host entry points");
            ST_TRY(ensure_rank_labels(h));
            db_feat = h->feat_rank.p;
        }
    }
    // a count matrix assembled by the caller (schic_mh_set_counts_device: blocks of a sharded symmetric K3) replaces K3
    const uint16_t* given = h->given_counts;
    const int64_t given_ld = h->given_ld;
    h->given_counts = nullptr;
    if (given && h->coll_blocks_open) reset_stage(h, {ST_SELECT, ST_RERANK, ST_GRAPH, ST_D2H, ST_PREP});
    else reset_stage(h, {ST_COLL, ST_SELECT, ST_RERANK, ST_GRAPH, ST_D2H, ST_PREP});
    h->coll_blocks_open = false;
    cudaStream_t s = h->stream;
    const uint32_t* q_sig = h->sig.p;
    if (!given && ndb >= h->p.max_bin_size && !packed_counts_possible(h, ndb)) {   // rule 5 (the packed K3 prunes while renaming)
        StageTimer t(h, ST_COLL, s);
        ST_TRY(prune_buckets(h, db_sig, ndb));
        db_sig = h->sig_masked.p;
        q_sig = h->sig_masked.p + row0 * (int64_t)H;
    }

    ST_TRY(h->cand_idx.reserve(nq * kcand, 0, s));
    ST_TRY(h->cand_cnt.reserve(nq * kcand, 0, s));
    ST_TRY(h->cand_nv.reserve(nq, 0, s));
    ST_TRY(h->sel_flags.reserve(nq, 0, s));
    if (!d_idx) {
        ST_TRY(h->out_idx.reserve(nq * k, 0, s));
        ST_TRY(h->out_dist.reserve(nq * k, 0, s));
        ST_TRY(h->out_nv.reserve(nq, 0, s));
        d_idx = h->out_idx.p; d_dist = h->out_dist.p; d_nv = h->out_nv.p;
    }

    // query blocking: the uint16 count matrix is produced and consumed in row blocks of <= 2 GiB
    const int64_t ld = given ? given_ld : ((ndb + 7) & ~(int64_t)7);
    // budget: a quarter of the free HBM, at least 2 GiB, at most 32 GiB (the all-rows-in-one-block case also
    // lets K3 use the symmetry of the counts)
    if (const char* forced = getenv("SCHIC_COUNT_BUDGET")) {     // tests: force several query blocks on small inputs
        h->count_budget = std::max<int64_t>(1, atoll(forced));
    } else if (h->count_budget == 0) {
        size_t free_b = 0, total_b = 0;
        h->count_budget = (int64_t)1 << 30;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            h->count_budget = std::max<int64_t>(h->count_budget, std::min<int64_t>((int64_t)(free_b / 4) / 2, (int64_t)1 << 34));
        else
            cudaGetLastError();
    }
    const int64_t budget_elems = getenv("SCHIC_COUNT_BUDGET") ? h->count_budget
                                                              : std::max<int64_t>(h->count_budget, h->counts.cap);   // never below what is already held
    int64_t qb = std::max<int64_t>(128, budget_elems / ld);
    qb = std::min<int64_t>((qb / 128) * 128, (nq + 127) / 128 * 128);
    if (given) qb = nq;
    else ST_TRY(h->counts.reserve(qb * ld, 0, s));
    int packed = 0;
    if (!given) {
        StageTimer t(h, ST_COLL, s);
        packed = prepare_packed_counts(h, db_sig, ndb);
        if (packed < 0) return packed;
    }
    const uint16_t* counts_in = given ? given : h->counts.p;
    const int64_t Hp = schic::k3h_padded_hashes(H);
    for (int64_t q0 = 0; q0 < nq; q0 += qb) {
        const int64_t q1 = std::min(nq, q0 + qb);
        if (!given) {
            StageTimer t(h, ST_COLL, s);
            if (packed)
                add_launches(h, launch_k3_packed(h, row0 + q0, q1 - q0, 0, ndb, h->counts.p, ld, nullptr, 0));
            else
                add_launches(h, schic::launch_collision_counts(q_sig + q0 * (int64_t)H, q1 - q0, db_sig, ndb, H,
                                                              h->counts.p, ld, s));
        }
        {
            StageTimer t(h, ST_SELECT, s);
            int nl = force_longrows() ? -1
                                      : schic::launch_select_topk(counts_in, q1 - q0, ndb, ld, H, h->p.minimal_blocks_in_common,
                                                                  kcand, h->cand_idx.p + q0 * kcand, h->cand_cnt.p + q0 * kcand,
                                                                  h->cand_nv.p + q0, h->sel_flags.p, s);
            if (nl < 0) {      // lists too long for the shared-memory select (e.g. the tool's default k = n): global sort
                int64_t per = 0;
                ST_TRY(reserve_long_keys(h, ndb, q1 - q0, &per));
                nl = schic::launch_select_topk_long(counts_in, q1 - q0, ndb, ld, H, h->p.minimal_blocks_in_common, kcand,
                                                    h->cand_idx.p + q0 * kcand, h->cand_cnt.p + q0 * kcand, h->cand_nv.p + q0,
                                                    h->long_keys.p, per, s);
            }
            add_launches(h, nl);
        }
        CU_TRY(cudaGetLastError());
    }
    if (fast) {
        StageTimer t(h, ST_SELECT, s);
        add_launches(h, schic::launch_fast_distance(h->cand_idx.p, h->cand_cnt.p, h->cand_nv.p, nq, kcand, k, H,
                                                   h->p.absolute_numbers, d_idx, d_dist, d_nv, s));
    } else {
        if (h->ext && h->db_ready) CU_TRY(cudaStreamWaitEvent(s, h->db_ready, 0));   // external CSR still in flight?
        const bool ext_prep = h->ext && h->ext_norm2 && h->ext_dir && h->ext_nw > 0;
        if (!ext_prep) {
            StageTimer t(h, ST_PREP, s);     // row norms + window directory of the database (once per fit)
            ST_TRY(ensure_norms(h, db_indptr, db_feat, db_val, ndb, h->ext ? -1 : (h->borrowed ? h->b_end : h->nnz)));
        }
        ST_TRY(rerank_lists(h, db_indptr, db_feat, db_val, ndb, ext_prep, row0, nq, h->cand_idx.p, h->cand_nv.p, kcand, k,
                            false, d_idx, d_dist, d_nv));
    }
    CU_TRY(cudaGetLastError());
    return 0;
}

// brute-force exact euclidean kNN (the scHicCluster -drm knn step) = K4 with every database row as a
// candidate of every query. Results in h->out_* unless device outputs are given.
int exact_knn_on_device(schic_mh_handle* h, int k, bool include_self, int32_t* d_idx, float* d_dist, int32_t* d_nv) {
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (k < 1) return fail(SCHIC_ERR_INVALID_ARG, "k must be >= 1");
    const int64_t nq = h->n;
    const int64_t ndb = h->ext ? h->db_n : h->n;
    const int64_t row0 = h->ext ? h->db_row0 : 0;
    const int64_t* db_indptr = h->ext ? h->db_indptr : h->d_indptr();
    const uint32_t* db_feat_in = h->ext ? h->db_feat : h->d_feat();
    const uint32_t* db_feat = db_feat_in;      // (replaced by the ids' ranks when ids need more than 32 bits)
    const float* db_val = h->ext ? h->db_val : h->d_val();
    const bool packed_db = h->ext && h->ext_packed && h->ext_flags && h->ext_norm2 && h->ext_dir && h->ext_nw > 0;
    if ((!db_val || !db_feat) && !packed_db) return fail(SCHIC_ERR_INVALID_ARG, "exact kNN needs the CSR values (fit with data)");
    if (h->have_hi) {
        if (h->ext || h->borrowed) return fail(SCHIC_ERR_UNSUPPORTED, "exact kNN with ids >= 2^32 needs the rows fitted through the host entry points");
        ST_TRY(ensure_rank_labels(h));
        db_feat = h->feat_rank.p;
    }
    if (ndb > INT32_MAX) return fail(SCHIC_ERR_UNSUPPORTED, "too many rows");
    reset_stage(h, {ST_COLL, ST_SELECT, ST_RERANK, ST_GRAPH, ST_D2H, ST_PREP});
    cudaStream_t s = h->stream;
    if (!d_idx) {
        ST_TRY(h->out_idx.reserve(nq * k, 0, s));
        ST_TRY(h->out_dist.reserve(nq * k, 0, s));
        ST_TRY(h->out_nv.reserve(nq, 0, s));
        d_idx = h->out_idx.p; d_dist = h->out_dist.p; d_nv = h->out_nv.p;
    }
    if (h->ext && h->db_ready) CU_TRY(cudaStreamWaitEvent(s, h->db_ready, 0));
    const bool ext_prep = h->ext && h->ext_norm2 && h->ext_dir && h->ext_nw > 0;
    if (!ext_prep) {
        StageTimer t(h, ST_PREP, s);
        ST_TRY(ensure_norms(h, db_indptr, db_feat, db_val, ndb, h->ext ? -1 : (h->borrowed ? h->b_end : h->nnz)));
    }
    return rerank_lists(h, db_indptr, db_feat, db_val, ndb, ext_prep, row0, nq, nullptr, nullptr, (int)ndb, k, !include_self,
                        d_idx, d_dist, d_nv);
}

// D2H of h->out_* ([n, ke]) into caller arrays of row stride k (columns [ke, k) padded).
int results_to_host(schic_mh_handle* h, int ke, int k, int32_t* idx, float* dist, int32_t* n_valid) {
    const int64_t n = h->n;
    {
        StageTimer t(h, ST_D2H, h->stream);
        if (ke == k) {
            CU_TRY(cudaMemcpyAsync(idx, h->out_idx.p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaMemcpyAsync(dist, h->out_dist.p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, h->stream));
        } else {   // k > n_fitted: pad columns [ke, k)
            for (int64_t e = 0; e < n * (int64_t)k; ++e) { idx[e] = -1; dist[e] = __builtin_inff(); }
            CU_TRY(cudaMemcpy2DAsync(idx, (size_t)k * 4, h->out_idx.p, (size_t)ke * 4, (size_t)ke * 4, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaMemcpy2DAsync(dist, (size_t)k * 4, h->out_dist.p, (size_t)ke * 4, (size_t)ke * 4, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        }
        if (n_valid) CU_TRY(cudaMemcpyAsync(n_valid, h->out_nv.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CU_TRY(cudaStreamSynchronize(h->stream));
    return check_range_flag(h);
}

// Graph emit from h->out_* ([n, ke]) and D2H of the CSR.
int graph_to_host(schic_mh_handle* h, int ke, int64_t* indptr, int32_t* indices, float* data) {
    const int64_t n = h->n;
    cudaStream_t s = h->stream;
    ST_TRY(h->g_indptr.reserve(n + 1, 0, s));
    ST_TRY(h->g_indices.reserve(n * ke, 0, s));
    ST_TRY(h->g_data.reserve(n * ke, 0, s));
    {
        StageTimer t(h, ST_GRAPH, s);
        int nl = force_longrows() ? -1 : schic::launch_graph_emit(h->out_idx.p, h->out_dist.p, h->out_nv.p, n, ke, h->g_indptr.p,
                                                                  h->g_indices.p, h->g_data.p, s);
        if (nl < 0) {          // rows longer than the shared-memory sort (dense graphs of > 16 384 cells): global sort
            int64_t per = 0;
            ST_TRY(reserve_long_keys(h, ke, n, &per));
            nl = schic::launch_graph_emit_long(h->out_idx.p, h->out_dist.p, h->out_nv.p, n, ke, h->g_indptr.p, h->g_indices.p,
                                               h->g_data.p, h->long_keys.p, per, s);
        }
        add_launches(h, nl);
    }
    CU_TRY(cudaGetLastError());
    {
        StageTimer t(h, ST_D2H, s);
        CU_TRY(cudaMemcpyAsync(indptr, h->g_indptr.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        const int64_t nnz = indptr[n];
        if (nnz > 0) {
            CU_TRY(cudaMemcpyAsync(indices, h->g_indices.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaMemcpyAsync(data, h->g_data.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, s));
        }
    }
    CU_TRY(cudaStreamSynchronize(s));
    return check_range_flag(h);
}

int effective_k(const schic_mh_handle* h, int k) {
    const int64_t ndb = h->ext ? h->db_n : h->n;
    return (int)std::min<int64_t>(k, ndb);
}

}  // namespace

extern "C" {

const char* schic_mh_last_error(void) { return g_err.c_str(); }
const char* schic_mh_backend(void) { return "cuda"; }

int schic_device_count(int32_t* n_out) {
    if (!n_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *n_out = n;
    return SCHIC_OK;
}

int schic_mh_create(const schic_mh_params* p, schic_mh_handle** out) {
    if (!p || !out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    if (p->struct_size != (int32_t)sizeof(schic_mh_params))
        return fail(SCHIC_ERR_INVALID_ARG, "schic_mh_params.struct_size mismatch");
    if (p->number_of_hash_functions < 1 || p->number_of_hash_functions > 65535)
        return fail(SCHIC_ERR_INVALID_ARG, "number_of_hash_functions must be in [1, 65535]");
    if (p->hash_width != 32 && p->hash_width != 64) return fail(SCHIC_ERR_INVALID_ARG, "hash_width must be 32 or 64");
    if (p->shingle != 0 && (p->shingle_size < 1 || p->number_of_hash_functions / p->shingle_size < 1))
        return fail(SCHIC_ERR_INVALID_ARG, "block combine needs 1 <= shingle_size <= number_of_hash_functions");
    if (p->n_neighbors < 1) return fail(SCHIC_ERR_INVALID_ARG, "n_neighbors must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SCHIC_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    int dev = p->device;
    if (dev < 0) CU_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(SCHIC_ERR_INVALID_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(SCHIC_ERR_NO_DEVICE, std::string("kernels are built for sm_100a only; device is ") + prop.name);
    CU_TRY(cudaSetDevice(dev));
    auto* h = new schic_mh_handle();
    h->p = *p;
    if (h->p.excess_factor < 1) h->p.excess_factor = 1;
    if (h->p.max_bin_size < 1) h->p.max_bin_size = INT64_MAX;
    if (p->shingle != 0 && p->shingle_size > 1) h->block = p->shingle_size;
    h->device = dev;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(SCHIC_ERR_CUDA, "cudaStreamCreate failed");
    }
    h->stream = h->own_stream;
    *out = h;
    return SCHIC_OK;
}

int schic_mh_destroy(schic_mh_handle* h) {
    if (!h) return SCHIC_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (cudaEvent_t e : h->pool) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    cudaGetLastError();
    return SCHIC_OK;
}

int schic_mh_set_stream(schic_mh_handle* h, void* cuda_stream) {
    ST_TRY(check(h));
    h->stream = (cudaStream_t)cuda_stream;
    return SCHIC_OK;
}

int schic_mh_synchronize(schic_mh_handle* h) {
    ST_TRY(check(h));
    ST_TRY(use_device(h));
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaStreamSynchronize(h->copy_stream));
    return check_range_flag(h);
}

int schic_mh_set_feature_range(schic_mh_handle* h, uint64_t n_features) {
    ST_TRY(check(h));
    h->feature_range = n_features;
    h->norm_of = nullptr;      // the window directory depends on it
    return SCHIC_OK;
}

// Values to the device on the copy stream: float32 as they are; uint16 / uint8 counts (a quarter to a half of the
// PCIe bytes -- Hi-C values are integer counts) into a staging buffer, widened to float32 by a kernel behind them.
static int upload_values(schic_mh_handle* h, const void* data, int32_t kind, size_t vbytes, int64_t first, int64_t pos0,
                         int64_t nnz, cudaStream_t cs) {
    const unsigned char* src = (const unsigned char*)data + (size_t)first * vbytes;
    if (kind == 0) {
        CU_TRY(cudaMemcpyAsync(h->val.p + pos0, src, (size_t)nnz * 4, cudaMemcpyHostToDevice, cs));
        return 0;
    }
    ST_TRY(h->val_stage.reserve((int64_t)((size_t)nnz * vbytes), 0, cs));
    CU_TRY(cudaMemcpyAsync(h->val_stage.p, src, (size_t)nnz * vbytes, cudaMemcpyHostToDevice, cs));
    add_launches(h, schic::launch_widen_counts(h->val_stage.p, kind == 1 ? 16 : 8, nnz, h->val.p + pos0, cs));
    CU_TRY(cudaGetLastError());
    return 0;
}

// data_kind: 0 = float32 values, 1 = uint16 counts, 2 = uint8 counts (expanded to float32 on the device)
// One staged value chunk (wire kind chosen by the worker that converted it) to the device behind the ids.
static int upload_value_chunk(schic_mh_handle* h, const Stager& st, int64_t c, int64_t pos0, cudaStream_t cs) {
    const int64_t a = st.chunk_begin(c), n = st.chunk_end(c) - a;
    const unsigned char* src = st.s_val + 4 * (size_t)a;
    const int kind = st.val_kind[c];
    if (kind == 0) {
        CU_TRY(cudaMemcpyAsync(h->val.p + pos0 + a, src, (size_t)n * 4, cudaMemcpyHostToDevice, cs));
        return 0;
    }
    unsigned char* dst = h->val_stage.p + 2 * (size_t)a;      // (2 bytes per non-zero reserved: chunks never overlap)
    CU_TRY(cudaMemcpyAsync(dst, src, (size_t)n * (kind == 1 ? 2 : 1), cudaMemcpyHostToDevice, cs));
    add_launches(h, schic::launch_widen_counts(dst, kind == 1 ? 16 : 8, n, h->val.p + pos0 + a, cs));
    CU_TRY(cudaGetLastError());
    return 0;
}

// data_kind: 0 = float32 values, 1 = uint16 counts, 2 = uint8 counts (expanded to float32 on the device).
// st != nullptr: indices / data are staged by st's worker threads (schic_mh_fit_csr_host); its data_kind may also be 3
// (float64) and the wire kind is chosen per chunk.
static int fit_csr_impl(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                        int32_t index_bits, const void* data, int32_t data_kind, int32_t append, bool async_upload,
                        Stager* st = nullptr) {
    if (!st && (data_kind < 0 || data_kind > 2)) return fail(SCHIC_ERR_INVALID_ARG, "data_kind must be 0 (f32), 1 (u16) or 2 (u8)");
    const size_t vbytes = data_kind == 0 ? 4 : (data_kind == 1 ? 2 : 1);
    ST_TRY(check(h));
    if (n_rows < 0 || !indptr) return fail(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
    if (index_bits != 32 && index_bits != 64) return fail(SCHIC_ERR_INVALID_ARG, "index_bits must be 32 or 64");
    const int64_t nnz = indptr[n_rows] - indptr[0];
    if (nnz < 0 || (nnz > 0 && !indices)) return fail(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
    if (!h->p.fast && nnz > 0 && !data) return fail(SCHIC_ERR_INVALID_ARG, "data is required when fast=0");
    for (int64_t r = 0; r < n_rows; ++r)
        if (indptr[r + 1] < indptr[r]) return fail(SCHIC_ERR_INVALID_ARG, "indptr not monotone");
    ST_TRY(use_device(h));
    if (h->borrowed && append) return fail(SCHIC_ERR_UNSUPPORTED, "cannot append to a borrowed device CSR");
    if (h->sliced) { CU_TRY(cudaStreamSynchronize(h->stream)); restore_fitted_signatures(h); }
    CU_TRY(cudaStreamSynchronize(h->copy_stream));   // an earlier async upload may still use the buffers / host vectors
    if (!append || h->borrowed) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->n = 0; h->nnz = 0; h->borrowed = false; h->have_val = false; h->have_hi = false;
        h->norm_of = nullptr; h->ext = false; h->db_ready = nullptr;
    }
    reset_stage(h, {ST_H2D, ST_SIG});
    h->k1_table_state = 0;
    h->sig16_of = nullptr; h->given_counts = nullptr; h->rank_valid = false;
    for (cudaEvent_t e : h->misc_ev) h->free_ev.push_back(e);
    h->misc_ev.clear();
    h->val_done = nullptr;
    h->norm_of = nullptr;
    const int H = h->p.number_of_hash_functions;
    const int64_t r0 = h->n, pos0 = h->nnz;
    const bool want_val = data != nullptr;
    if (r0 > 0 && want_val != h->have_val) return fail(SCHIC_ERR_INVALID_ARG, "partial_fit must be consistent about data");
    cudaStream_t s = h->stream, cs = h->copy_stream;

    // 32-bit feature ids (width 32 hashes (uint32)(f+1), so truncation is exact there); the width-64 hash also needs
    // the high words, but only when some id really is >= 2^32 -- 64-bit arrays of small ids stay on the 32-bit paths
    const uint32_t* f32 = nullptr;
    std::vector<uint32_t>& lo = h->host_lo;
    std::vector<uint32_t>& hi = h->host_hi;
    bool any_hi = false;
    if (st) {
        // staged: the workers fill page-locked buffers in upload order while this thread enqueues what is ready
        ST_TRY(h->pin_ids.reserve((size_t)std::max<int64_t>(nnz, 1) * 4));
        if (data) ST_TRY(h->pin_val.reserve((size_t)std::max<int64_t>(nnz, 1) * 4));
        st->indices = indices; st->index_bits = index_bits; st->data = data; st->data_kind = data ? data_kind : -1;
        st->first = indptr[0]; st->nnz = nnz;
        st->s_ids = (uint32_t*)h->pin_ids.p; st->s_val = h->pin_val.p;
        int nt = h->p.number_of_cores > 0 ? h->p.number_of_cores : (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("SCHIC_STAGE_THREADS")) nt = atoi(e);
        st->start(std::max(1, std::min(nt, 32)));
        f32 = st->s_ids;
    } else if (index_bits == 32) {
        f32 = (const uint32_t*)indices + indptr[0];
    } else {
        const uint64_t* f64 = (const uint64_t*)indices + indptr[0];
        lo.resize(nnz);
        for (int64_t e = 0; e < nnz; ++e) {
            lo[e] = (uint32_t)f64[e];
            any_hi = any_hi || (f64[e] >> 32) != 0;
        }
        f32 = lo.data();
    }
    // high words are kept from the first batch that needs them on (earlier batches are back-filled with zeros)
    // (the width-64 hash needs them; so does the exact rerank, which runs on the ranks of the full ids: k7_relabel.cu)
    const bool want_hi = !st && (h->p.hash_width == 64 || !h->p.fast) && (any_hi || (r0 > 0 && h->have_hi));
    if (st && r0 > 0 && h->have_hi) return fail(SCHIC_ERR_UNSUPPORTED, "staged fit cannot append to rows with ids >= 2^32");
    if (want_hi) {
        hi.assign(nnz, 0u);
        if (any_hi) {
            const uint64_t* f64 = (const uint64_t*)indices + indptr[0];
            for (int64_t e = 0; e < nnz; ++e) hi[e] = (uint32_t)(f64[e] >> 32);
        }
    }

    ST_TRY(h->indptr.reserve(r0 + n_rows + 1, r0 + 1, s));
    ST_TRY(h->feat.reserve(pos0 + nnz, pos0, s));
    if (want_val) ST_TRY(h->val.reserve(pos0 + nnz, pos0, s));
    if (want_hi) {
        const bool backfill = r0 > 0 && !h->have_hi;
        ST_TRY(h->feat_hi.reserve(pos0 + nnz, backfill ? 0 : pos0, s));
        if (backfill && pos0 > 0) CU_TRY(cudaMemsetAsync(h->feat_hi.p, 0, (size_t)pos0 * sizeof(uint32_t), s));
    }
    ST_TRY(h->sig.reserve((r0 + n_rows) * (int64_t)H, r0 * (int64_t)H, s));

    // rebased indptr (absolute positions in the device store)
    std::vector<int64_t> ip(n_rows + 1);
    for (int64_t r = 0; r <= n_rows; ++r) ip[r] = pos0 + (indptr[r] - indptr[0]);

    cudaEvent_t up0 = get_event(h), up1 = get_event(h);
    if (up0) cudaEventRecord(up0, cs);
    CU_TRY(cudaMemcpyAsync(h->indptr.p + r0, ip.data(), (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
    CU_TRY(cudaStreamSynchronize(cs));   // ip is a local vector
    h->n = r0 + n_rows;
    h->nnz = pos0 + nnz;
    h->have_val = want_val;
    h->have_hi = want_hi;
    if (!st) h->fit_data_kind = (r0 > 0 && h->fit_data_kind != data_kind) ? 0 : data_kind;   // appended batches: the weakest guarantee

    // Upload the feature ids in row-aligned chunks on the copy stream; K1 for chunk i runs on the
    // compute stream while chunk i+1 is still in flight.
    // A small first chunk (8 M non-zeros, ~0.6 ms of PCIe) lets K1 start early; later chunks are 32 M so that
    // every K1 launch still holds several waves of row work items.
    cudaEvent_t sig0 = get_event(h);
    bool sig_started = false, val_enqueued = false;
    int64_t c0 = 0;
    // (a batch that may take the shared-table K1 path is hashed in one piece after its last chunk has landed:
    // the table is built from all of its feature ids)
    const bool whole_batch = k1_table_candidate(h, n_rows, nnz);
    // With the caller's feature-range hint the table does not depend on the data at all (it is built for every id
    // of the range): it is computed on the compute stream while the ids are still crossing PCIe, and the batch only
    // needs the lookup + fallback once it has landed.
    h->k1t_prebuilt = false;
    // (only when building for the WHOLE range costs less than the upload it hides behind: ~2.4e12 table evaluations/s
    // against ~55 GB/s of ids, i.e. range * H < 175 * nnz; otherwise the table is built from the ids present, later)
    if (whole_batch && k1_mode_from_env() == K1_AUTO && h->feature_range > 0 && h->feature_range <= 134217728ull &&
        (double)h->feature_range * (double)H < 175.0 * (double)nnz &&
        schic::k1t_applicable(n_rows, nnz, (uint32_t)(h->feature_range - 1), H)) {
        const uint32_t max_id = (uint32_t)(h->feature_range - 1);
        const size_t ecap = schic::k1t_entries_capacity(n_rows, nnz, max_id, H);
        ST_TRY(ensure_counters(h));
        if (k1t_cache_hit(h, max_id, n_rows, nnz)) {
            h->k1_table_state = 3;                    // the whole-range table of an earlier fit is still there
        } else {
            h->k1t_cache.valid = false;
            ST_TRY(h->k1t_present.reserve((int64_t)max_id + 1, 0, s));
            ST_TRY(h->k1t_slot.reserve((int64_t)max_id + 1, 0, s));
            ST_TRY(h->k1t_entries.reserve((int64_t)ecap, 0, s));
            const int nl = schic::launch_signatures_table(h->p.hash_width, nullptr, nullptr, n_rows, nnz, 0, max_id, H, nullptr,
                                                          h->k1t_present.p, reinterpret_cast<uint2*>(h->k1t_slot.p),
                                                          h->k1t_entries.p, ecap, h->k1t_counters.p, nullptr, 1, s);
            if (nl < 0) return fail(SCHIC_ERR_CUDA, "K1 table build failed");
            add_launches(h, nl);
            CU_TRY(cudaGetLastError());
            k1t_cache_store(h, max_id, n_rows, nnz);
            h->k1_table_state = 2;
        }
        h->k1t_prebuilt = true;
    }
    while (c0 < n_rows) {
        const int64_t kChunkNnz = (whole_batch && !h->k1t_prebuilt) ? INT64_MAX : (int64_t)(c0 == 0 ? 8 : 32) << 20;
        int64_t c1 = c0 + 1;
        while (c1 < n_rows && (ip[c1 + 1] - ip[c0]) <= kChunkNnz) ++c1;
        const int64_t a = ip[c0] - pos0, b = ip[c1] - pos0;
        if (st) {
            st->wait_ids_upto(b);
            if (st->any_hi.load()) return fail(SCHIC_ERR_UNSUPPORTED, "feature ids >= 2^32: use schic_mh_fit_csr");
        }
        if (b > a) {
            CU_TRY(cudaMemcpyAsync(h->feat.p + pos0 + a, f32 + a, (size_t)(b - a) * 4, cudaMemcpyHostToDevice, cs));
            if (want_hi)
                CU_TRY(cudaMemcpyAsync(h->feat_hi.p + pos0 + a, hi.data() + a, (size_t)(b - a) * 4, cudaMemcpyHostToDevice, cs));
        }
        cudaEvent_t ready = get_misc_event(h);
        if (!ready) return fail(SCHIC_ERR_CUDA, "cudaEventCreate failed");
        CU_TRY(cudaEventRecord(ready, cs));
        if (c1 == n_rows && want_val && nnz > 0 && !st) {
            // the values follow the last feature chunk at once, so the copy engine stays busy while K1 runs
            // (and while the host may block on K1's id-range read-back)
            ST_TRY(upload_values(h, data, data_kind, vbytes, indptr[0], pos0, nnz, cs));
            val_enqueued = true;
        }
        CU_TRY(cudaStreamWaitEvent(s, ready, 0));
        if (!sig_started && sig0) { cudaEventRecord(sig0, s); sig_started = true; }
        ST_TRY(hash_rows(h, r0 + c0, r0 + c1, ip[c0], ip[c1], s));
        c0 = c1;
    }
    h->k1t_prebuilt = false;      // valid for exactly this batch
    cudaEvent_t sig1 = get_event(h);
    if (sig_started && sig1) { cudaEventRecord(sig1, s); h->ev[ST_SIG].push_back({sig0, sig1}); }
    if (want_val && nnz > 0 && st) {
        // staged values: chunk by chunk as the workers finish them (K1 of the last id chunk is already queued)
        ST_TRY(h->val_stage.reserve(2 * nnz, 0, cs));
        int weakest = 2;
        for (int64_t c = 0; c < st->n_chunks; ++c) {
            Stager::wait(st->val_done[c]);
            ST_TRY(upload_value_chunk(h, *st, c, pos0, cs));
            weakest = std::min(weakest, (int)st->val_kind[c]);
        }
        h->fit_data_kind = (r0 > 0 && h->fit_data_kind != weakest) ? 0 : weakest;
        val_enqueued = true;
    }
    if (st) st->join();
    if (want_val && nnz > 0 && !val_enqueued) ST_TRY(upload_values(h, data, data_kind, vbytes, indptr[0], pos0, nnz, cs));
    if (up0 && up1) { cudaEventRecord(up1, cs); h->ev[ST_H2D].push_back({up0, up1}); }
    // the host vectors (lo/hi) and the caller's arrays must outlive the copies; later stages use
    // the compute stream, so make it wait for the value upload as well
    cudaEvent_t done = get_misc_event(h);
    if (!done) return fail(SCHIC_ERR_CUDA, "cudaEventCreate failed");
    CU_TRY(cudaEventRecord(done, cs));
    if (async_upload) {
        h->val_done = done;                   // the rerank preparation waits for it (ensure_norms)
    } else {
        CU_TRY(cudaStreamWaitEvent(s, done, 0));
        CU_TRY(cudaStreamSynchronize(cs));
    }
    return apply_block_combine(h);
}

int schic_mh_fit_csr(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                     int32_t index_bits, const float* data, int32_t append) {
    return fit_csr_impl(h, n_rows, indptr, indices, index_bits, data, 0, append, false);
}

int schic_mh_fit_csr_async(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                           int32_t index_bits, const float* data, int32_t append) {
    return fit_csr_impl(h, n_rows, indptr, indices, index_bits, data, 0, append, true);
}

int schic_mh_fit_csr_counts(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                            int32_t index_bits, const void* counts, int32_t count_bits, int32_t append,
                            int32_t async_upload) {
    if (count_bits != 8 && count_bits != 16) return fail(SCHIC_ERR_INVALID_ARG, "count_bits must be 8 or 16");
    return fit_csr_impl(h, n_rows, indptr, indices, index_bits, counts, count_bits == 16 ? 1 : 2, append, async_upload != 0);
}

static int fit_csr_device_impl(schic_mh_handle* h, int64_t n_rows, int64_t nnz, int64_t pos0, const int64_t* d_indptr,
                               const uint32_t* d_indices, const float* d_data, int32_t append) {
    ST_TRY(check(h));
    if (append) return fail(SCHIC_ERR_UNSUPPORTED, "device-resident fit does not append");
    if (n_rows < 0 || nnz < 0 || !d_indptr || (nnz > 0 && !d_indices)) return fail(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
    if (!h->p.fast && !d_data) return fail(SCHIC_ERR_INVALID_ARG, "data is required when fast=0");
    ST_TRY(use_device(h));
    if (h->sliced) { CU_TRY(cudaStreamSynchronize(h->stream)); restore_fitted_signatures(h); }
    reset_stage(h, {ST_H2D, ST_SIG});
    h->k1_table_state = 0;
    h->sig16_of = nullptr; h->given_counts = nullptr;
    const int H = h->p.number_of_hash_functions;
    h->k1t_prebuilt = false;
    h->borrowed = true; h->b_indptr = d_indptr; h->b_feat = d_indices; h->b_val = d_data;
    h->have_val = d_data != nullptr; h->have_hi = false;
    h->n = n_rows; h->nnz = nnz; h->norm_of = nullptr; h->ext = false; h->db_ready = nullptr;
    ST_TRY(h->sig.reserve(n_rows * (int64_t)H, 0, h->stream));
    if (pos0 < 0) {
        // indptr[0] of a borrowed CSR not given: read back one value (8 bytes, drains the stream) -- positions are absolute
        CU_TRY(cudaMemcpyAsync(&pos0, d_indptr, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    h->b_end = pos0 + nnz;
    h->fit_data_kind = 0;
    {
        StageTimer t(h, ST_SIG, h->stream);
        ST_TRY(hash_rows(h, 0, n_rows, pos0, pos0 + nnz, h->stream));
    }
    return apply_block_combine(h);
}

int schic_mh_fit_csr_host(schic_mh_handle* h, int64_t n_rows, const void* indptr, int32_t indptr_bits, const void* indices,
                          int32_t index_bits, const void* data, int32_t data_kind, int32_t append, int32_t async_upload) {
    ST_TRY(check(h));
    if (n_rows < 0 || !indptr) return fail(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
    if (indptr_bits != 32 && indptr_bits != 64) return fail(SCHIC_ERR_INVALID_ARG, "indptr_bits must be 32 or 64");
    if (data && (data_kind < 0 || data_kind > 3)) return fail(SCHIC_ERR_INVALID_ARG, "data_kind must be 0 (f32), 1 (u16), 2 (u8) or 3 (f64)");
    std::vector<int64_t> ip64;
    const int64_t* ip = (const int64_t*)indptr;
    if (indptr_bits == 32) {
        ip64.resize(n_rows + 1);
        for (int64_t r = 0; r <= n_rows; ++r) ip64[r] = ((const int32_t*)indptr)[r];
        ip = ip64.data();
    }
    Stager st;
    return fit_csr_impl(h, n_rows, ip, indices, index_bits, data, data_kind, append, async_upload != 0, &st);
}

int schic_mh_fit_csr_device(schic_mh_handle* h, int64_t n_rows, int64_t nnz, const int64_t* d_indptr,
                            const uint32_t* d_indices, const float* d_data, int32_t append) {
    return fit_csr_device_impl(h, n_rows, nnz, -1, d_indptr, d_indices, d_data, append);
}

int schic_mh_fit_csr_device_at(schic_mh_handle* h, int64_t n_rows, int64_t nnz, int64_t indptr0, const int64_t* d_indptr,
                               const uint32_t* d_indices, const float* d_data, int32_t append) {
    if (indptr0 < 0) return fail(SCHIC_ERR_INVALID_ARG, "indptr0 must be the (non-negative) value of d_indptr[0]");
    return fit_csr_device_impl(h, n_rows, nnz, indptr0, d_indptr, d_indices, d_data, append);
}

int schic_mh_n_fitted(const schic_mh_handle* h, int64_t* n_out) {
    ST_TRY(check(h));
    if (!n_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *n_out = h->n;
    return SCHIC_OK;
}

int schic_mh_signatures(schic_mh_handle* h, uint32_t* out) {
    ST_TRY(check(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (!out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    ST_TRY(use_device(h));
    CU_TRY(cudaMemcpyAsync(out, h->sig.p, (size_t)h->n * h->p.number_of_hash_functions * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return check_range_flag(h);
}

int schic_mh_signatures_device(schic_mh_handle* h, void** d_sig_out, int64_t* n_rows_out) {
    ST_TRY(check(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (!d_sig_out || !n_rows_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *d_sig_out = h->sig.p;
    *n_rows_out = h->n;
    return SCHIC_OK;
}

int schic_mh_csr_device(schic_mh_handle* h, const int64_t** d_indptr_out, const uint32_t** d_indices_out,
                        const float** d_data_out, int64_t* nnz_out) {
    ST_TRY(check(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (!d_indptr_out || !d_indices_out || !d_data_out || !nnz_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *d_indptr_out = h->d_indptr();
    *d_indices_out = h->d_feat();
    *d_data_out = h->d_val();
    *nnz_out = h->nnz;
    return SCHIC_OK;
}

int schic_mh_set_database_device(schic_mh_handle* h, int64_t n_total, int64_t row0, const uint32_t* d_sig_all,
                                 const int64_t* d_indptr_all, const uint32_t* d_indices_all, const float* d_data_all) {
    ST_TRY(check(h));
    h->db_ready = nullptr;
    h->ext_norm2 = nullptr; h->ext_dir = nullptr; h->ext_nw = 0;
    h->ext_packed = nullptr; h->ext_flags = nullptr; h->ext_nnz = 0;
    h->sig16_of = nullptr; h->given_counts = nullptr;
    if (!d_sig_all) { h->ext = false; h->norm_of = nullptr; return SCHIC_OK; }
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "fit the local shard first");
    if (h->sliced) return fail(SCHIC_ERR_UNSUPPORTED, "external database while an active-hash prefix or block combine (shingle) is set");
    if (row0 < 0 || row0 + h->n > n_total) return fail(SCHIC_ERR_INVALID_ARG, "local rows do not fit in the database");
    h->ext = true; h->db_n = n_total; h->db_row0 = row0; h->db_sig = d_sig_all;
    h->db_indptr = d_indptr_all; h->db_feat = d_indices_all; h->db_val = d_data_all;
    h->norm_of = nullptr;
    return SCHIC_OK;
}

int schic_mh_set_database_ready_event(schic_mh_handle* h, void* cuda_event) {
    ST_TRY(check(h));
    h->db_ready = (cudaEvent_t)cuda_event;
    return SCHIC_OK;
}

int schic_mh_rerank_windows(schic_mh_handle* h, int32_t* n_windows_out) {
    ST_TRY(check(h));
    if (!n_windows_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *n_windows_out = 0;
    const char* mode = getenv("SCHIC_K4_MODE");
    if (mode && (strcmp(mode, "hash") == 0 || strcmp(mode, "window64") == 0)) return SCHIC_OK;
    if (h->feature_range == 0 || h->feature_range > 0xFFFFFFFFull) return SCHIC_OK;
    const int nw = schic::k4c_window_count((uint32_t)(h->feature_range - 1));
    if (nw <= 8192) *n_windows_out = nw;
    return SCHIC_OK;
}

int schic_mh_rerank_prep_rows(schic_mh_handle* h, int64_t n_rows, const int64_t* d_indptr, const uint32_t* d_indices,
                              const float* d_data, int32_t n_windows, double* d_norm2_out, uint32_t* d_dir_out,
                              uint32_t* d_packed_out, uint32_t* d_flags_out) {
    ST_TRY(check(h));
    if (n_rows < 0 || !d_indptr || !d_norm2_out || !d_dir_out || !d_flags_out || n_windows < 1)
        return fail(SCHIC_ERR_INVALID_ARG, "bad arguments");
    if (n_rows > 0 && (!d_indices || !d_data)) return fail(SCHIC_ERR_INVALID_ARG, "the rerank preparation needs ids and values");
    ST_TRY(use_device(h));
    CU_TRY(cudaMemsetAsync(d_flags_out, 0, 4 * sizeof(uint32_t), h->stream));
    add_launches(h, schic::launch_rerank_prep(d_indptr, d_indices, d_data, n_rows, n_windows, d_norm2_out, d_dir_out,
                                              d_packed_out, d_flags_out, h->stream));
    CU_TRY(cudaGetLastError());
    return SCHIC_OK;
}

int schic_mh_set_database_prep_device(schic_mh_handle* h, const double* d_norm2_all, const uint32_t* d_dir_all,
                                      int32_t n_windows, const uint32_t* d_packed_all, const uint32_t* d_flags_all,
                                      int64_t nnz_total) {
    ST_TRY(check(h));
    if (!h->ext) return fail(SCHIC_ERR_INVALID_ARG, "set the external database first");
    if (d_packed_all && ((uintptr_t)d_packed_all & 15u)) return fail(SCHIC_ERR_INVALID_ARG, "the packed stream must be 16-byte aligned");
    if (d_packed_all && (!d_flags_all || nnz_total < 0)) return fail(SCHIC_ERR_INVALID_ARG, "a packed stream needs its flags and length");
    h->ext_norm2 = d_norm2_all; h->ext_dir = d_dir_all; h->ext_nw = (d_norm2_all && d_dir_all) ? n_windows : 0;
    h->ext_packed = h->ext_nw > 0 ? d_packed_all : nullptr;
    h->ext_flags = h->ext_packed ? d_flags_all : nullptr;
    h->ext_nnz = h->ext_packed ? nnz_total : 0;
    return SCHIC_OK;
}

int schic_mh_rerank_kernel_name(schic_mh_handle* h, char* out, int32_t n_out) {
    ST_TRY(check(h));
    if (!out || n_out < 32) return fail(SCHIC_ERR_INVALID_ARG, "out must hold 32 bytes");
    ST_TRY(use_device(h));
    const char* name = "none";
    if (h->rerank_path == 3) {
        name = h->n_windows > 0 || (h->ext && h->ext_nw > 0) ? "k4_rerank_window_kernel" : "k4_rerank_kernel";
    } else if (h->rerank_path == 1) {
        name = schic::k4c_kernel_name();
    } else if (h->rerank_path == 2) {      // both were launched: the flags on the device say which one did the work
        const uint32_t* fl = (h->ext && h->ext_flags) ? h->ext_flags : h->k4_flags.p;
        const int64_t nnz = (h->ext && h->ext_flags) ? h->ext_nnz : h->prep_nnz;
        uint32_t f[4] = {0, 0, 0, 0};
        CU_TRY(cudaMemcpyAsync(f, fl, sizeof(f), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        name = (f[0] == 0 && f[1] <= esc_limit_for(nnz)) ? schic::k4c_kernel_name() : "k4_rerank_window_kernel";
    }
    snprintf(out, (size_t)n_out, "%s", name);
    return SCHIC_OK;
}

int schic_mh_collision_counts(schic_mh_handle* h, int64_t row0, int64_t row1, uint16_t* out) {
    ST_TRY(check(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (row0 < 0 || row1 > h->n || row0 > row1 || !out) return fail(SCHIC_ERR_INVALID_ARG, "bad row range");
    ST_TRY(use_device(h));
    const int H = h->p.number_of_hash_functions;
    const int64_t ndb = h->ext ? h->db_n : h->n;
    const uint32_t* db_sig = h->ext ? h->db_sig : h->sig.p;
    const uint32_t* q_sig = h->sig.p;
    if (ndb >= h->p.max_bin_size && !packed_counts_possible(h, ndb)) {
        ST_TRY(prune_buckets(h, db_sig, ndb));
        db_sig = h->sig_masked.p;
        q_sig = h->sig_masked.p + (h->ext ? h->db_row0 : 0) * (int64_t)H;
    }
    const int64_t ld = (ndb + 7) & ~(int64_t)7;
    const int64_t nq = row1 - row0;
    if (nq == 0) return SCHIC_OK;
    ST_TRY(h->counts.reserve(nq * ld, 0, h->stream));
    const int packed = prepare_packed_counts(h, db_sig, ndb);
    if (packed < 0) return packed;
    if (packed) {
        const int64_t first = (h->ext ? h->db_row0 : 0) + row0;     // the query rows are database rows
        add_launches(h, launch_k3_packed(h, first, nq, 0, ndb, h->counts.p, ld, nullptr, 0));
    } else {
        add_launches(h, schic::launch_collision_counts(q_sig + row0 * (int64_t)H, nq, db_sig, ndb, H, h->counts.p, ld, h->stream));
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy2DAsync(out, (size_t)ndb * 2, h->counts.p, (size_t)ld * 2, (size_t)ndb * 2, (size_t)nq,
                             cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return SCHIC_OK;
}

int schic_mh_collision_block_device(schic_mh_handle* h, int64_t q_row0, int64_t q_rows, int64_t d_row0, int64_t d_rows,
                                    uint16_t* d_cnt, int64_t ld, uint16_t* d_cnt_t, int64_t ld_t) {
    ST_TRY(check(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    ST_TRY(use_device(h));
    const int H = h->p.number_of_hash_functions;
    const int64_t ndb = h->ext ? h->db_n : h->n;
    const uint32_t* db_sig = h->ext ? h->db_sig : h->sig.p;
    if (q_row0 < 0 || d_row0 < 0 || q_rows < 0 || d_rows < 0 || q_row0 + q_rows > ndb || d_row0 + d_rows > ndb || !d_cnt || ld < d_rows ||
        (d_cnt_t && ld_t < q_rows))
        return fail(SCHIC_ERR_INVALID_ARG, "bad block arguments");
    if (!packed_counts_possible(h, ndb)) return fail(SCHIC_ERR_UNSUPPORTED, "the block entry point needs the packed K3 (H <= 4094, <= 57344 database rows)");
    cudaStream_t s = h->stream;
    if (!h->coll_blocks_open) { reset_stage(h, {ST_COLL}); h->coll_blocks_open = true; }
    StageTimer t(h, ST_COLL, s);
    if (h->sig16_of != (const void*)db_sig || h->sig16_rows != ndb) {      // rename once per database
        const int r = prepare_packed_counts(h, db_sig, ndb);
        if (r <= 0) return r < 0 ? r : fail(SCHIC_ERR_UNSUPPORTED, "packed K3 not applicable");
        h->sig16_of = db_sig; h->sig16_rows = ndb;
    }
    if (q_rows == 0 || d_rows == 0) return SCHIC_OK;
    const bool same = q_row0 == d_row0 && q_rows == d_rows;          // same rows twice: symmetric kernel, no transposed copy
    add_launches(h, launch_k3_packed(h, q_row0, q_rows, d_row0, d_rows, d_cnt, ld, same ? nullptr : d_cnt_t, ld_t));
    CU_TRY(cudaGetLastError());
    return SCHIC_OK;
}

int schic_mh_set_counts_device(schic_mh_handle* h, const uint16_t* d_counts, int64_t ld) {
    ST_TRY(check(h));
    const int64_t ndb = h->ext ? h->db_n : h->n;
    if (d_counts && (ld < ndb || (ld & 7) != 0 || ((uintptr_t)d_counts & 15u)))
        return fail(SCHIC_ERR_INVALID_ARG, "the count matrix needs a 16-byte aligned base and a row stride >= database rows that is a multiple of 8");
    h->given_counts = d_counts;
    h->given_ld = ld;
    return SCHIC_OK;
}

int schic_mh_kneighbors_device(schic_mh_handle* h, int32_t k, int32_t fast, int32_t* d_idx, float* d_dist, int32_t* d_nvalid) {
    ST_TRY(check(h));
    if (!d_idx || !d_dist || !d_nvalid) return fail(SCHIC_ERR_INVALID_ARG, "null output");
    ST_TRY(use_device(h));
    if (h->n > 0 && k > effective_k(h, k)) return fail(SCHIC_ERR_INVALID_ARG, "k exceeds the number of fitted rows");
    return kneighbors_on_device(h, k, fast, d_idx, d_dist, d_nvalid);
}

int schic_mh_kneighbors(schic_mh_handle* h, int32_t k, int32_t fast, int32_t* idx, float* dist, int32_t* n_valid) {
    ST_TRY(check(h));
    if (k < 1 || !idx || !dist) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors arguments");
    ST_TRY(use_device(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    const int ke = effective_k(h, k);
    ST_TRY(kneighbors_on_device(h, ke, fast, nullptr, nullptr, nullptr));
    return results_to_host(h, ke, k, idx, dist, n_valid);
}

int schic_mh_kneighbors_graph(schic_mh_handle* h, int32_t k, int32_t fast, int64_t* indptr, int32_t* indices, float* data) {
    ST_TRY(check(h));
    if (k < 1 || !indptr || !indices || !data) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors_graph arguments");
    ST_TRY(use_device(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    const int ke = effective_k(h, k);
    ST_TRY(kneighbors_on_device(h, ke, fast, nullptr, nullptr, nullptr));
    return graph_to_host(h, ke, indptr, indices, data);
}

int schic_mh_kneighbors_graph_device(schic_mh_handle* h, int32_t k, int32_t fast, int64_t* d_indptr, int32_t* d_indices,
                                     float* d_data) {
    ST_TRY(check(h));
    if (k < 1 || !d_indptr || !d_indices || !d_data) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors_graph arguments");
    ST_TRY(use_device(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    const int ke = effective_k(h, k);
    ST_TRY(kneighbors_on_device(h, ke, fast, nullptr, nullptr, nullptr));
    StageTimer t(h, ST_GRAPH, h->stream);
    int nl = force_longrows() ? -1 : schic::launch_graph_emit(h->out_idx.p, h->out_dist.p, h->out_nv.p, h->n, ke, d_indptr,
                                                              d_indices, d_data, h->stream);
    if (nl < 0) {
        int64_t per = 0;
        ST_TRY(reserve_long_keys(h, ke, h->n, &per));
        nl = schic::launch_graph_emit_long(h->out_idx.p, h->out_dist.p, h->out_nv.p, h->n, ke, d_indptr, d_indices, d_data,
                                           h->long_keys.p, per, h->stream);
    }
    add_launches(h, nl);
    CU_TRY(cudaGetLastError());
    return SCHIC_OK;
}

int schic_mh_kneighbors_exact(schic_mh_handle* h, int32_t k, int32_t include_self, int32_t* idx, float* dist, int32_t* n_valid) {
    ST_TRY(check(h));
    if (k < 1 || !idx || !dist) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors arguments");
    ST_TRY(use_device(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    const int ke = std::max(1, std::min(k, effective_k(h, INT32_MAX) - (include_self ? 0 : 1)));
    ST_TRY(exact_knn_on_device(h, ke, include_self != 0, nullptr, nullptr, nullptr));
    return results_to_host(h, ke, k, idx, dist, n_valid);
}

int schic_mh_kneighbors_exact_graph(schic_mh_handle* h, int32_t k, int32_t include_self, int64_t* indptr, int32_t* indices, float* data) {
    ST_TRY(check(h));
    if (k < 1 || !indptr || !indices || !data) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors_graph arguments");
    ST_TRY(use_device(h));
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    const int ke = std::max(1, std::min(k, effective_k(h, INT32_MAX) - (include_self ? 0 : 1)));
    ST_TRY(exact_knn_on_device(h, ke, include_self != 0, nullptr, nullptr, nullptr));
    return graph_to_host(h, ke, indptr, indices, data);
}

int schic_mh_kneighbors_exact_device(schic_mh_handle* h, int32_t k, int32_t include_self, int32_t* d_idx, float* d_dist, int32_t* d_nvalid) {
    ST_TRY(check(h));
    if (!d_idx || !d_dist || !d_nvalid) return fail(SCHIC_ERR_INVALID_ARG, "null output");
    ST_TRY(use_device(h));
    if (h->n > 0 && k > effective_k(h, k)) return fail(SCHIC_ERR_INVALID_ARG, "k exceeds the number of fitted rows");
    return exact_knn_on_device(h, k, include_self != 0, d_idx, d_dist, d_nvalid);
}

int schic_mh_set_active_hashes(schic_mh_handle* h, int32_t n_hashes) {
    ST_TRY(check(h));
    ST_TRY(use_device(h));
    if (h->block > 1) return fail(SCHIC_ERR_UNSUPPORTED, "active-hash prefix with block combine (shingle)");
    const int H_fit = h->sliced ? h->H_fit : h->p.number_of_hash_functions;
    if (n_hashes < 1 || n_hashes > H_fit) return fail(SCHIC_ERR_INVALID_ARG, "n_hashes must be in [1, number_of_hash_functions as created]");
    if (h->ext) return fail(SCHIC_ERR_UNSUPPORTED, "active-hash prefix with an external database");
    cudaStream_t s = h->stream;
    CU_TRY(cudaStreamSynchronize(s));
    restore_fitted_signatures(h);
    if (n_hashes == H_fit) return SCHIC_OK;
    h->H_fit = H_fit;
    std::swap(h->sig.p, h->sig_fit.p);     // sig_fit: fitted matrix; sig: the compacted prefix (buffer reused between calls)
    std::swap(h->sig.cap, h->sig_fit.cap);
    h->sliced = true;
    h->p.number_of_hash_functions = n_hashes;
    ST_TRY(h->sig.reserve(std::max<int64_t>(1, h->n * (int64_t)n_hashes), 0, s));
    if (h->n > 0)
        CU_TRY(cudaMemcpy2DAsync(h->sig.p, (size_t)n_hashes * 4, h->sig_fit.p, (size_t)H_fit * 4, (size_t)n_hashes * 4,
                                 (size_t)h->n, cudaMemcpyDeviceToDevice, s));
    return SCHIC_OK;
}

int schic_mh_stage_ms(schic_mh_handle* h, float* out, int32_t n_out) {
    ST_TRY(check(h));
    if (!out || n_out < ST_COUNT + 1) return fail(SCHIC_ERR_INVALID_ARG, "out must hold 9 floats");
    ST_TRY(use_device(h));
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaStreamSynchronize(h->copy_stream));
    float sum = 0.f;
    for (int st = 0; st < ST_COUNT; ++st) {
        float acc = 0.f;
        for (auto& pr : h->ev[st]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) acc += ms; else cudaGetLastError();
        }
        out[st] = acc;
        sum += acc;
    }
    out[ST_COUNT] = sum;
    return SCHIC_OK;
}

int schic_mh_k1_table_state(schic_mh_handle* h, int32_t* state_out) {
    ST_TRY(check(h));
    if (!state_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *state_out = h->k1_table_state;
    return SCHIC_OK;
}

int schic_mh_launch_count(schic_mh_handle* h, int64_t* n_out) {
    ST_TRY(check(h));
    if (!n_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *n_out = h->launches;
    return SCHIC_OK;
}

int schic_int32_peak(int32_t device, int32_t iters, double* out) {
    if (!out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SCHIC_ERR_NO_DEVICE, "no CUDA device"); }
    if (device >= 0) CU_TRY(cudaSetDevice(device));
    for (int i = 0; i < 16; ++i) out[i] = 0.0;
    if (schic::run_int32_peak(iters < 64 ? 64 : iters, out, nullptr) < 0) return fail(SCHIC_ERR_CUDA, "int32 peak kernel failed");
    return SCHIC_OK;
}

int schic_k1_variant_ms(int32_t variant, int64_t n_rows, int64_t nnz, const int64_t* d_indptr, const uint32_t* d_indices,
                        int32_t n_hash, uint32_t* d_sig, int32_t reps, float* ms_out) {
    if (!d_indptr || !d_indices || !d_sig || !ms_out || reps < 1) return fail(SCHIC_ERR_INVALID_ARG, "bad arguments");
    int64_t pos0 = 0;
    CU_TRY(cudaMemcpy(&pos0, d_indptr, sizeof(int64_t), cudaMemcpyDeviceToHost));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    void* scratch = nullptr;
    unsigned char* present = nullptr; uint2* slot = nullptr; uint16_t* entries = nullptr; uint32_t* counters = nullptr;
    uint32_t max_id = 0;
    size_t ecap = 0;
    if (variant >= 200) {          // shared-table path
        CU_TRY(cudaMalloc(&counters, 4 * sizeof(uint32_t)));
        CU_TRY(cudaMemset(counters, 0, 4 * sizeof(uint32_t)));
        schic::launch_max_id(d_indices + pos0, nnz, counters + 2, nullptr);
        CU_TRY(cudaMemcpy(&max_id, counters + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (max_id >= 134217728u) return fail(SCHIC_ERR_UNSUPPORTED, "id range too large for the table path");
        ecap = schic::k1t_entries_capacity(n_rows, nnz, max_id, n_hash);
        CU_TRY(cudaMalloc(&present, (size_t)max_id + 1));
        CU_TRY(cudaMalloc(&slot, ((size_t)max_id + 1) * sizeof(uint2)));
        CU_TRY(cudaMalloc(&entries, ecap * sizeof(uint16_t)));
        CU_TRY(cudaMalloc(&scratch, schic::k1t_items_bytes(n_rows, nnz)));
    } else if (variant >= 100) {
        CU_TRY(cudaMalloc(&scratch, schic::k1f_scratch_bytes(n_rows, nnz)));
    }
    auto launch = [&]() {
        if (variant >= 200)
            return schic::launch_signatures_table(32, d_indptr, d_indices, n_rows, nnz, pos0, max_id, n_hash, d_sig, present,
                                                    slot, entries, ecap, counters, scratch, 0, nullptr);
        return variant >= 100 ? schic::launch_signatures32_filter(d_indptr, d_indices, n_rows, nnz, n_hash, d_sig, scratch, variant == 101 ? 1 : 0, nullptr)
                              : schic::launch_signatures32_variant(variant, d_indptr, d_indices, n_rows, nnz, pos0, n_hash, d_sig, nullptr);
    };
    if (launch() < 0) return fail(SCHIC_ERR_UNSUPPORTED, "variant not compiled in");
    CU_TRY(cudaEventRecord(e0, nullptr));
    for (int r = 0; r < reps; ++r) launch();
    CU_TRY(cudaEventRecord(e1, nullptr));
    CU_TRY(cudaEventSynchronize(e1));
    CU_TRY(cudaEventElapsedTime(ms_out, e0, e1));
    *ms_out /= (float)reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (variant >= 200) {
        uint32_t c2[2] = {0, 0};
        cudaMemcpy(c2, counters, sizeof(c2), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[k1 table] max_id %u entries %u of %zu overflow %u\n", max_id, c2[0], ecap, c2[1]);
    }
    if (scratch) cudaFree(scratch);
    if (present) cudaFree(present);
    if (slot) cudaFree(slot);
    if (entries) cudaFree(entries);
    if (counters) cudaFree(counters);
    CU_TRY(cudaGetLastError());
    return SCHIC_OK;
}

int schic_pinned_alloc(int64_t bytes, void** out) {
    if (!out || bytes < 0) return fail(SCHIC_ERR_INVALID_ARG, "bad arguments");
    CU_TRY(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return SCHIC_OK;
}

int schic_pinned_free(void* p) {
    if (p) CU_TRY(cudaFreeHost(p));
    return SCHIC_OK;
}

}  // extern "C"
