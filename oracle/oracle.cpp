// oracle.cpp -- CPU restatement (C++17 + OpenMP) of the MinHash approximate-kNN path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may load this library. The product path
// (schicexplorer_b200/csrc) never links, loads or falls back to it.
//
// PARITY UNPINNED. The arithmetic of this path lives in the third-party package
// `sparse-neighbors-search >= 0.7` (/root/reference/requirements.txt:3, setup.py:96), which is
// neither vendored in /root/reference nor installable here, and the reference's own tests
// (schicexplorer/test/test_scHicClusterMinHash.py:35-63) assert only the cell-name column and
// "3 distinct labels". This file therefore restates the published algorithm as fixed by
// SURVEY.md section 8(c) (the normative spec, rules 1-11), anchored on:
//   - docs/content/tools/scHicClusterMinHash.rst:11   "32-bit mix function by Thomas Wang"
//   - docs/content/Example_analysis.rst:464-467       per sample and hash function: hash every
//                                                     non-zero feature id, keep the minimum;
//                                                     similarity = number of hash collisions
//   - schicexplorer/utilities.py:119-120              feature id = bin1 * nbins + bin2
//   - schicexplorer/scHicClusterMinHash.py:347-348, 402-404   the kwargs
//   - schicexplorer/scHicClusterMinHash.py:355-357            kneighbors_graph(mode='distance')
// It exports the C ABI of include/schic_mh.h so tests swap back ends by dlopen path.
//
// The code is written the way the upstream CPU path works (inverse index of hash buckets,
// per-query collision counting through the buckets, merge-join euclidean distances), which makes
// it an independent check of the GPU's all-pairs formulation rather than a copy of it.

#include "../include/schic_mh.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

constexpr uint32_t M = SCHIC_MH_MODULUS;

// Rule 2, width 32: h_j(f) = wang32((f+1)*(j+1)) mod M, all arithmetic uint32 with wraparound.
inline uint32_t wang32(uint32_t k) {
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057u;
    k ^= k >> 16;
    return k;
}
inline uint32_t hash32(uint64_t f, uint32_t j) {
    return wang32((uint32_t)(f + 1) * (j + 1)) % M;
}
// Rule 2, width 64 (alternative HashSpec): same formula evaluated in uint64 before the modulo.
inline uint32_t hash64(uint64_t f, uint32_t j) {
    uint64_t k = (f + 1) * (uint64_t)(j + 1);
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057u;
    k ^= k >> 16;
    return (uint32_t)(k % M);
}

// Rule 4 (block combine, "shingle"; decision point D2 -- OFF unless schic_mh_params.shingle != 0): the minima of
// `shingle_size` consecutive hash functions are folded into ONE key, key = sig[b*s]; key = h'(sig[b*s+t] + 1, key + 1)
// for t = 1..s-1, which leaves H / s tables. h' is the hash family's own mix applied to the product of its two
// arguments (SURVEY leaves h' open; this is the restatement's choice, recorded with the HashSpec).
inline uint32_t combine32(uint32_t v, uint32_t key) { return wang32((v + 1u) * (key + 1u)) % M; }
inline uint32_t combine64(uint32_t v, uint32_t key) {
    uint64_t k = ((uint64_t)v + 1) * ((uint64_t)key + 1);
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057u;
    k ^= k >> 16;
    return (uint32_t)(k % M);
}

}  // namespace

struct schic_mh_handle {
    schic_mh_params p;
    int64_t n = 0;
    std::vector<int64_t> indptr{0};
    std::vector<uint64_t> idx;
    std::vector<float> val;
    std::vector<uint32_t> sig;  // [n, H]
    int H_created = 0;          // table count the handle was created with (schic_mh_set_active_hashes changes p's)
    int H_raw = 0, block = 1;   // hash functions evaluated per row; minima folded per table (1: block combine off)
    // inverse index (rule 5), rebuilt lazily after every fit
    bool index_ok = false;
    std::vector<int32_t> order;        // [H, n] cells of table j sorted by (value, cell)
    std::vector<int32_t> bucket_begin; // [H, n] per cell: start of its bucket in order[j]
    std::vector<int32_t> bucket_end;   // [H, n]
};

namespace {

int nthreads(const schic_mh_handle* h) {
#ifdef _OPENMP
    return h->p.number_of_cores > 0 ? h->p.number_of_cores : omp_get_max_threads();
#else
    (void)h;
    return 1;
#endif
}

// Inner loops of rule 2-3 for one row. target_clones gives an AVX2 and an AVX-512 body with
// runtime dispatch, so one prebuilt .so is safe on whatever host CPU the GPU box has.
__attribute__((target_clones("default", "avx2", "arch=x86-64-v4")))
void sig_row32(const uint64_t* f, int64_t nnz, int H, uint32_t* s) {
    for (int64_t e = 0; e < nnz; ++e) {
        const uint32_t f1 = (uint32_t)(f[e] + 1);
        for (int j = 0; j < H; ++j) {
            const uint32_t v = wang32(f1 * (uint32_t)(j + 1)) % M;
            s[j] = v < s[j] ? v : s[j];
        }
    }
}
__attribute__((target_clones("default", "avx2", "arch=x86-64-v4")))
void sig_row64(const uint64_t* f, int64_t nnz, int H, uint32_t* s) {
    for (int64_t e = 0; e < nnz; ++e) {
        for (int j = 0; j < H; ++j) {
            const uint32_t v = hash64(f[e], (uint32_t)j);
            s[j] = v < s[j] ? v : s[j];
        }
    }
}

// Rules 2-3: signature of the rows [r0, r1) of the handle's CSR.
void compute_signatures(schic_mh_handle* h, int64_t r0, int64_t r1) {
    const int H = h->p.number_of_hash_functions;     // tables = columns of sig
    const bool w64 = h->p.hash_width == 64;
    h->sig.resize((size_t)r1 * H);
    if (h->block <= 1) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads(h))
        for (int64_t c = r0; c < r1; ++c) {
            uint32_t* s = h->sig.data() + (size_t)c * H;
            for (int j = 0; j < H; ++j) s[j] = M;  // empty row => sentinel M
            const uint64_t* f = h->idx.data() + h->indptr[c];
            const int64_t nnz = h->indptr[c + 1] - h->indptr[c];
            if (w64) sig_row64(f, nnz, H, s); else sig_row32(f, nnz, H, s);
        }
        return;
    }
    // rule 4: H_raw minima per row, folded block by block into H keys; an empty row keeps the sentinel
    const int Hr = h->H_raw, bs = h->block;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads(h))
    for (int64_t c = r0; c < r1; ++c) {
        std::vector<uint32_t> raw((size_t)Hr, M);
        const uint64_t* f = h->idx.data() + h->indptr[c];
        const int64_t nnz = h->indptr[c + 1] - h->indptr[c];
        if (w64) sig_row64(f, nnz, Hr, raw.data()); else sig_row32(f, nnz, Hr, raw.data());
        uint32_t* s = h->sig.data() + (size_t)c * H;
        for (int b = 0; b < H; ++b) {
            uint32_t key = raw[(size_t)b * bs];
            if (nnz > 0)
                for (int t = 1; t < bs; ++t)
                    key = w64 ? combine64(raw[(size_t)b * bs + t], key) : combine32(raw[(size_t)b * bs + t], key);
            s[b] = nnz > 0 ? key : M;
        }
    }
}

inline bool admissible_value(uint32_t v) { return v != 0u && v != M; }  // rule 5

// Rule 5: per table j the buckets of equal signature value.
void build_index(schic_mh_handle* h) {
    if (h->index_ok) return;
    const int H = h->p.number_of_hash_functions;
    const int64_t n = h->n;
    h->order.assign((size_t)H * n, 0);
    h->bucket_begin.assign((size_t)H * n, 0);
    h->bucket_end.assign((size_t)H * n, 0);
#pragma omp parallel for schedule(static) num_threads(nthreads(h))
    for (int j = 0; j < H; ++j) {
        int32_t* ord = h->order.data() + (size_t)j * n;
        for (int64_t c = 0; c < n; ++c) ord[c] = (int32_t)c;
        const uint32_t* sig = h->sig.data();
        std::sort(ord, ord + n, [&](int32_t a, int32_t b) {
            const uint32_t va = sig[(size_t)a * H + j], vb = sig[(size_t)b * H + j];
            return va != vb ? va < vb : a < b;
        });
        int32_t* bb = h->bucket_begin.data() + (size_t)j * n;
        int32_t* be = h->bucket_end.data() + (size_t)j * n;
        int64_t s = 0;
        while (s < n) {
            int64_t e = s + 1;
            const uint32_t v = sig[(size_t)ord[s] * H + j];
            while (e < n && sig[(size_t)ord[e] * H + j] == v) ++e;
            for (int64_t q = s; q < e; ++q) {
                bb[ord[q]] = (int32_t)s;
                be[ord[q]] = (int32_t)e;
            }
            s = e;
        }
    }
    h->index_ok = true;
}

// Rules 5-6: collision counts of query row c against every fitted row, through the buckets.
void count_row(const schic_mh_handle* h, int64_t c, uint16_t* cnt /* [n], zeroed */) {
    const int H = h->p.number_of_hash_functions;
    const int64_t n = h->n;
    for (int j = 0; j < H; ++j) {
        const uint32_t v = h->sig[(size_t)c * H + j];
        if (!admissible_value(v)) continue;
        const int32_t b = h->bucket_begin[(size_t)j * n + c], e = h->bucket_end[(size_t)j * n + c];
        if ((int64_t)(e - b) >= h->p.max_bin_size) continue;
        const int32_t* ord = h->order.data() + (size_t)j * n;
        for (int32_t q = b; q < e; ++q) ++cnt[ord[q]];
    }
}

// Rule 9: euclidean distance over the union of supports; fp64 accumulation, rounded to fp32.
float euclid(const schic_mh_handle* h, int64_t a, int64_t b) {
    int64_t i = h->indptr[a], ie = h->indptr[a + 1];
    int64_t j = h->indptr[b], je = h->indptr[b + 1];
    double acc = 0.0;
    while (i < ie && j < je) {
        if (h->idx[i] == h->idx[j]) {
            const double d = (double)h->val[i] - (double)h->val[j];
            acc += d * d; ++i; ++j;
        } else if (h->idx[i] < h->idx[j]) {
            acc += (double)h->val[i] * (double)h->val[i]; ++i;
        } else {
            acc += (double)h->val[j] * (double)h->val[j]; ++j;
        }
    }
    for (; i < ie; ++i) acc += (double)h->val[i] * (double)h->val[i];
    for (; j < je; ++j) acc += (double)h->val[j] * (double)h->val[j];
    return (float)std::sqrt(acc);
}

struct Cand { int32_t idx; int32_t cnt; float dist; };

// Rules 7-9 for one query row. Returns the number of valid neighbours written (<= k).
int neighbours_of(const schic_mh_handle* h, int64_t c, int k, bool fast, std::vector<uint16_t>& cnt,
                  std::vector<Cand>& cand, int32_t* idx_out, float* dist_out) {
    const int64_t n = h->n;
    const int H = h->p.number_of_hash_functions;
    std::fill(cnt.begin(), cnt.end(), (uint16_t)0);
    count_row(h, c, cnt.data());
    cand.clear();
    for (int64_t d = 0; d < n; ++d)
        if ((int)cnt[d] >= h->p.minimal_blocks_in_common && cnt[d] > 0)
            cand.push_back({(int32_t)d, (int32_t)cnt[d], 0.f});
    std::sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) {
        return a.cnt != b.cnt ? a.cnt > b.cnt : a.idx < b.idx;
    });
    const size_t keep = (size_t)k * (size_t)std::max(1, h->p.excess_factor);
    if (cand.size() > keep) cand.resize(keep);
    if (fast) {
        for (auto& x : cand)
            x.dist = h->p.absolute_numbers ? (float)x.cnt : 1.0f - (float)x.cnt / (float)H;  // rule 8
    } else {
        for (auto& x : cand) x.dist = euclid(h, c, x.idx);
        std::sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) {
            return a.dist != b.dist ? a.dist < b.dist : a.idx < b.idx;
        });
    }
    if (cand.size() > (size_t)k) cand.resize(k);
    for (int t = 0; t < k; ++t) {
        if ((size_t)t < cand.size()) { idx_out[t] = cand[t].idx; dist_out[t] = cand[t].dist; }
        else { idx_out[t] = -1; dist_out[t] = std::numeric_limits<float>::infinity(); }
    }
    return (int)cand.size();
}

int check_handle(const schic_mh_handle* h) {
    if (!h) return fail(SCHIC_ERR_INVALID_ARG, "null handle");
    return 0;
}

}  // namespace

extern "C" {

const char* schic_mh_last_error(void) { return g_err.c_str(); }
const char* schic_mh_backend(void) { return "oracle"; }

int schic_mh_create(const schic_mh_params* p, schic_mh_handle** out) {
    if (!p || !out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    if (p->struct_size != (int32_t)sizeof(schic_mh_params))
        return fail(SCHIC_ERR_INVALID_ARG, "schic_mh_params.struct_size mismatch");
    if (p->number_of_hash_functions < 1 || p->number_of_hash_functions > 65535)
        return fail(SCHIC_ERR_INVALID_ARG, "number_of_hash_functions must be in [1, 65535]");
    if (p->hash_width != 32 && p->hash_width != 64)
        return fail(SCHIC_ERR_INVALID_ARG, "hash_width must be 32 or 64");
    if (p->shingle != 0 && (p->shingle_size < 1 || p->number_of_hash_functions / p->shingle_size < 1))
        return fail(SCHIC_ERR_INVALID_ARG, "block combine needs 1 <= shingle_size <= number_of_hash_functions");
    if (p->n_neighbors < 1) return fail(SCHIC_ERR_INVALID_ARG, "n_neighbors must be >= 1");
    auto* h = new schic_mh_handle();
    h->p = *p;
    h->H_raw = p->number_of_hash_functions;
    if (p->shingle != 0 && p->shingle_size > 1) {   // rule 4: everything behind the signatures sees H / s tables
        h->block = p->shingle_size;
        h->p.number_of_hash_functions = h->H_raw / h->block;
    }
    h->H_created = h->p.number_of_hash_functions;
    if (h->p.excess_factor < 1) h->p.excess_factor = 1;
    if (h->p.max_bin_size < 1) h->p.max_bin_size = std::numeric_limits<int64_t>::max();
    *out = h;
    return SCHIC_OK;
}

int schic_mh_destroy(schic_mh_handle* h) { delete h; return SCHIC_OK; }

int schic_mh_fit_csr(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                     int32_t index_bits, const float* data, int32_t append) {
    if (int e = check_handle(h)) return e;
    if (n_rows < 0 || !indptr) return fail(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
    if (index_bits != 32 && index_bits != 64) return fail(SCHIC_ERR_INVALID_ARG, "index_bits must be 32 or 64");
    const int64_t nnz = indptr[n_rows] - indptr[0];
    if (nnz < 0 || (nnz > 0 && !indices)) return fail(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
    if (!h->p.fast && nnz > 0 && !data) return fail(SCHIC_ERR_INVALID_ARG, "data is required when fast=0");
    if (!append) {
        h->n = 0; h->indptr.assign(1, 0); h->idx.clear(); h->val.clear(); h->sig.clear();
    }
    if (h->p.number_of_hash_functions != h->H_created) {   // a fit ends schic_mh_set_active_hashes
        h->p.number_of_hash_functions = h->H_created;
        h->sig.clear();
        compute_signatures(h, 0, h->n);
    }
    const int64_t r0 = h->n;
    const int64_t base = h->indptr.back();
    for (int64_t r = 0; r < n_rows; ++r) {
        if (indptr[r + 1] < indptr[r]) return fail(SCHIC_ERR_INVALID_ARG, "indptr not monotone");
        h->indptr.push_back(base + (indptr[r + 1] - indptr[0]));
    }
    h->idx.reserve(h->idx.size() + nnz);
    for (int64_t e = indptr[0]; e < indptr[n_rows]; ++e)
        h->idx.push_back(index_bits == 32 ? (uint64_t)((const uint32_t*)indices)[e] : ((const uint64_t*)indices)[e]);
    for (int64_t e = indptr[0]; e < indptr[n_rows]; ++e) h->val.push_back(data ? data[e] : 1.0f);
    h->n = r0 + n_rows;
    compute_signatures(h, r0, h->n);
    h->index_ok = false;
    return SCHIC_OK;
}

int schic_mh_n_fitted(const schic_mh_handle* h, int64_t* n_out) {
    if (int e = check_handle(h)) return e;
    if (!n_out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    *n_out = h->n;
    return SCHIC_OK;
}

int schic_mh_signatures(schic_mh_handle* h, uint32_t* out) {
    if (int e = check_handle(h)) return e;
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (!out) return fail(SCHIC_ERR_INVALID_ARG, "null argument");
    std::memcpy(out, h->sig.data(), h->sig.size() * sizeof(uint32_t));
    return SCHIC_OK;
}

int schic_mh_collision_counts(schic_mh_handle* h, int64_t row0, int64_t row1, uint16_t* out) {
    if (int e = check_handle(h)) return e;
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (row0 < 0 || row1 > h->n || row0 > row1 || !out) return fail(SCHIC_ERR_INVALID_ARG, "bad row range");
    build_index(h);
    const int64_t n = h->n;
    std::memset(out, 0, (size_t)(row1 - row0) * n * sizeof(uint16_t));
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads(h))
    for (int64_t c = row0; c < row1; ++c) count_row(h, c, out + (size_t)(c - row0) * n);
    return SCHIC_OK;
}

int schic_mh_kneighbors(schic_mh_handle* h, int32_t k, int32_t fast, int32_t* idx, float* dist, int32_t* n_valid) {
    if (int e = check_handle(h)) return e;
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (k < 1 || !idx || !dist) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors arguments");
    const bool is_fast = fast < 0 ? h->p.fast != 0 : fast != 0;
    build_index(h);
    const int64_t n = h->n;
#pragma omp parallel num_threads(nthreads(h))
    {
        std::vector<uint16_t> cnt(n);
        std::vector<Cand> cand;
#pragma omp for schedule(dynamic, 4)
        for (int64_t c = 0; c < n; ++c) {
            const int nv = neighbours_of(h, c, k, is_fast, cnt, cand, idx + (size_t)c * k, dist + (size_t)c * k);
            if (n_valid) n_valid[c] = nv;
        }
    }
    return SCHIC_OK;
}

// Rule 10: row c holds its neighbours sorted by column.
static void lists_to_graph(const schic_mh_handle* h, int k, const std::vector<int32_t>& idx, const std::vector<float>& dist,
                           const std::vector<int32_t>& nv, int64_t* indptr, int32_t* indices, float* data) {
    const int64_t n = h->n;
    indptr[0] = 0;
    for (int64_t c = 0; c < n; ++c) indptr[c + 1] = indptr[c] + nv[c];
#pragma omp parallel for schedule(static) num_threads(nthreads(h))
    for (int64_t c = 0; c < n; ++c) {
        std::vector<std::pair<int32_t, float>> row(nv[c]);
        for (int t = 0; t < nv[c]; ++t) row[t] = {idx[(size_t)c * k + t], dist[(size_t)c * k + t]};
        std::sort(row.begin(), row.end());
        for (int t = 0; t < nv[c]; ++t) {
            indices[indptr[c] + t] = row[t].first;
            data[indptr[c] + t] = row[t].second;
        }
    }
}

int schic_mh_kneighbors_graph(schic_mh_handle* h, int32_t k, int32_t fast, int64_t* indptr, int32_t* indices, float* data) {
    if (int e = check_handle(h)) return e;
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (k < 1 || !indptr || !indices || !data) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors_graph arguments");
    const int64_t n = h->n;
    std::vector<int32_t> idx((size_t)n * k), nv(n);
    std::vector<float> dist((size_t)n * k);
    if (int e = schic_mh_kneighbors(h, k, fast, idx.data(), dist.data(), nv.data())) return e;
    lists_to_graph(h, k, idx, dist, nv, indptr, indices, data);
    return SCHIC_OK;
}

int schic_mh_kneighbors_exact(schic_mh_handle* h, int32_t k, int32_t include_self, int32_t* idx, float* dist, int32_t* n_valid);

int schic_mh_kneighbors_exact_graph(schic_mh_handle* h, int32_t k, int32_t include_self, int64_t* indptr, int32_t* indices, float* data) {
    if (int e = check_handle(h)) return e;
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (k < 1 || !indptr || !indices || !data) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors_graph arguments");
    const int64_t n = h->n;
    std::vector<int32_t> idx((size_t)n * k), nv(n);
    std::vector<float> dist((size_t)n * k);
    if (int e = schic_mh_kneighbors_exact(h, k, include_self, idx.data(), dist.data(), nv.data())) return e;
    lists_to_graph(h, k, idx, dist, nv, indptr, indices, data);
    return SCHIC_OK;
}

// the exact-kNN sibling (scHicCluster.py:179-180 — sklearn NearestNeighbors(...).fit(X) followed by
// kneighbors_graph(mode='distance'), which on sparse input is brute-force euclidean and leaves the query itself out).
// Every fitted row is a candidate; order (dist asc, idx asc); keep k. include_self != 0 keeps the query row.
int schic_mh_kneighbors_exact(schic_mh_handle* h, int32_t k, int32_t include_self, int32_t* idx, float* dist, int32_t* n_valid) {
    if (int e = check_handle(h)) return e;
    if (h->n == 0) return fail(SCHIC_ERR_NOT_FITTED, "no rows fitted");
    if (k < 1 || !idx || !dist) return fail(SCHIC_ERR_INVALID_ARG, "bad kneighbors arguments");
    if (h->val.empty() && !h->idx.empty()) return fail(SCHIC_ERR_INVALID_ARG, "exact kNN needs the CSR values (fit with data)");
    const int64_t n = h->n;
#pragma omp parallel num_threads(nthreads(h))
    {
        std::vector<Cand> cand;
#pragma omp for schedule(dynamic, 4)
        for (int64_t c = 0; c < n; ++c) {
            cand.clear();
            for (int64_t d = 0; d < n; ++d)
                if (include_self || d != c) cand.push_back({(int32_t)d, 0, euclid(h, c, d)});
            std::sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) {
                return a.dist != b.dist ? a.dist < b.dist : a.idx < b.idx;
            });
            if (cand.size() > (size_t)k) cand.resize(k);
            for (int t = 0; t < k; ++t) {
                if ((size_t)t < cand.size()) { idx[(size_t)c * k + t] = cand[t].idx; dist[(size_t)c * k + t] = cand[t].dist; }
                else { idx[(size_t)c * k + t] = -1; dist[(size_t)c * k + t] = std::numeric_limits<float>::infinity(); }
            }
            if (n_valid) n_valid[c] = (int32_t)cand.size();
        }
    }
    return SCHIC_OK;
}

// SURVEY §8(f)-4 (hyperopt sweeps over -nh, scHicClusterMinHashHyperopt.py:119-148): from now on the handle behaves as
// if created with n_hashes hash functions. The checker does NOT use the prefix property of the hash family: it simply
// re-hashes every fitted row with n_hashes functions.
int schic_mh_set_active_hashes(schic_mh_handle* h, int32_t n_hashes) {
    if (int e = check_handle(h)) return e;
    if (h->block > 1) return fail(SCHIC_ERR_UNSUPPORTED, "active-hash prefix with block combine (shingle)");
    if (n_hashes < 1 || n_hashes > h->H_created)
        return fail(SCHIC_ERR_INVALID_ARG, "n_hashes must be in [1, number_of_hash_functions as created]");
    if (n_hashes == h->p.number_of_hash_functions) return SCHIC_OK;
    h->p.number_of_hash_functions = n_hashes;
    h->sig.clear();
    compute_signatures(h, 0, h->n);
    h->index_ok = false;
    return SCHIC_OK;
}

}  // extern "C"
