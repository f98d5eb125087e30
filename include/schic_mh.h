/*
 * schic_mh.h -- C ABI of the MinHash approximate-kNN hot path of scHicClusterMinHash.
 *
 * This is the drop-in boundary. The reference (joachimwolff/scHiCExplorer) reaches this path
 * through the Python classes `sparse_neighbors_search.MinHash` / `MinHashClustering`
 * (import: schicexplorer/scHicClusterMinHash.py:21-22); those classes are a thin Python shell
 * over a native extension. Every entry point below names the reference call site it serves.
 *
 * Two shared libraries export this exact symbol set:
 *   - schicexplorer_b200/csrc/libschic_mh.so   the product: CUDA sm_100a kernels, no CPU fallback
 *   - oracle/liboracle.so                      test infrastructure: CPU restatement (C++/OpenMP)
 * so the tests swap back ends by dlopen path.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * schic_status on failure; the message is available from schic_mh_last_error() (thread local).
 * No C++ exception and no sticky CUDA error crosses this boundary. A handle is bound to one
 * device and one stream; calls on one handle must be serialised by the caller.
 */
#ifndef SCHIC_MH_H_
#define SCHIC_MH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCHIC_MH_ABI_VERSION 1

/* Modulus of the hash family: signature values live in [0, SCHIC_MH_MODULUS);
 * SCHIC_MH_MODULUS itself is the "empty row" sentinel (SURVEY.md 8(c) rules 2-3). */
#define SCHIC_MH_MODULUS 2147483647u

typedef enum schic_status {
    SCHIC_OK = 0,
    SCHIC_ERR_INVALID_ARG = -1,
    SCHIC_ERR_NOT_FITTED = -2,
    SCHIC_ERR_NO_DEVICE = -3,   /* CUDA library only: no usable GPU / driver */
    SCHIC_ERR_CUDA = -4,        /* CUDA runtime error (message has the cudaError string) */
    SCHIC_ERR_OOM = -5,
    SCHIC_ERR_UNSUPPORTED = -6
} schic_status;

/* Mirrors the kwargs scHiCExplorer passes to MinHash(...)
 * (scHicClusterMinHash.py:347-348 saveMemory branch, :402-404 default branch). */
typedef struct schic_mh_params {
    int32_t struct_size;              /* = sizeof(schic_mh_params), ABI guard */
    int32_t n_neighbors;              /* n_neighbors=            (-k / ncells) */
    int32_t number_of_hash_functions; /* number_of_hash_functions= (-nh)        */
    int32_t hash_width;               /* 32 (default HashSpec) or 64; SURVEY 8(c) D1 */
    int32_t fast;                     /* fast=  1: collision distances, 0: exact euclidean rerank */
    int32_t absolute_numbers;         /* absolute_numbers= (always False in the reference) */
    int32_t minimal_blocks_in_common; /* minimal_blocks_in_common= (100 at :404) */
    int32_t excess_factor;            /* excess_factor= (1 at :404) */
    int64_t max_bin_size;             /* max_bin_size= (100000 at :403); buckets >= this are skipped */
    int32_t shingle_size;             /* minima folded per table when shingle != 0; otherwise accepted and unused (D2) */
    int32_t shingle;                  /* 0 (default HashSpec): one table per hash function. != 0: block combine (SURVEY 8(c)
                                         rule 4) -- K1 evaluates number_of_hash_functions minima per row, K2 folds every
                                         shingle_size consecutive ones into a key, and everything behind the signatures
                                         (signature accessors, counts, thresholds, 1 - cnt/H) sees
                                         number_of_hash_functions / shingle_size tables. Not available together with
                                         schic_mh_set_active_hashes or an external (sharded) database. */
    int32_t number_of_cores;          /* oracle: OpenMP threads (<=0: all); CUDA: ignored */
    int32_t device;                   /* CUDA ordinal, -1 = current device; oracle: ignored */
} schic_mh_params;

typedef struct schic_mh_handle schic_mh_handle;

/* MinHash.__init__ (scHicClusterMinHash.py:347, :402). */
int schic_mh_create(const schic_mh_params* params, schic_mh_handle** out);
int schic_mh_destroy(schic_mh_handle* h);
const char* schic_mh_last_error(void);
/* "cuda" or "oracle". */
const char* schic_mh_backend(void);

/* MinHash.fit(X) (append=0, :351, and inside MinHashClustering.fit :406) and
 * MinHash.partial_fit(X=...) (append=1, :353): hashes the rows of a host CSR matrix and appends
 * them to the fitted set; appended rows get ids n_fitted, n_fitted+1, ...
 *   indptr     [n_rows+1] int64, indptr[0] may be non-zero
 *   indices    feature ids bin1*nbins+bin2 (utilities.py:119-120), uint32 if index_bits==32,
 *              uint64 if index_bits==64; need not be sorted for hashing, MUST be sorted and
 *              duplicate free per row when fast==0 (scipy CSR built by the loader is)
 *   data       [nnz] float32 counts; may be NULL when the handle was created with fast=1 */
int schic_mh_fit_csr(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr,
                     const void* indices, int32_t index_bits, const float* data, int32_t append);

int schic_mh_n_fitted(const schic_mh_handle* h, int64_t* n_out);

/* Parity hook: the MinHash signature matrix, row major uint32 [n_fitted, H]. */
int schic_mh_signatures(schic_mh_handle* h, uint32_t* out);

/* Parity hook: collision counts cnt[c][d] for query rows c in [row0,row1) against all fitted
 * rows d, row major uint16 [(row1-row0), n_fitted] (SURVEY 8(c) rules 5-6). */
int schic_mh_collision_counts(schic_mh_handle* h, int64_t row0, int64_t row1, uint16_t* out);

/* MinHash.kneighbors(): for every fitted row the <= k neighbours (self included) in result
 * order (fast: count desc / index asc; exact: distance asc / index asc).
 *   fast  -1: as created, 0/1 override (the Python kneighbors(fast=...) kwarg)
 *   idx   [n,k] int32, padded with -1;  dist [n,k] float32, padded with +inf
 *   n_valid [n] int32 */
int schic_mh_kneighbors(schic_mh_handle* h, int32_t k, int32_t fast,
                        int32_t* idx, float* dist, int32_t* n_valid);

/* MinHash.kneighbors_graph(mode='distance') (scHicClusterMinHash.py:355): CSR n x n, row c holds
 * its <= k (neighbour, distance) pairs sorted by column. Caller provides indptr [n+1] int64,
 * indices [n*k] int32, data [n*k] float32; nnz = indptr[n]. */
int schic_mh_kneighbors_graph(schic_mh_handle* h, int32_t k, int32_t fast,
                              int64_t* indptr, int32_t* indices, float* data);

/* SURVEY 8(f)-3, the exact-kNN sibling: scHicCluster -drm knn (scHicCluster.py:179-180) runs sklearn
 * NearestNeighbors(n_neighbors=k, algorithm='ball_tree').fit(X).kneighbors_graph(mode='distance'), which on a sparse
 * X is brute-force euclidean kNN and leaves the query row itself out. Here: every fitted row is a candidate of every
 * fitted row (handle fitted WITH data), order (distance asc, index asc), keep k.
 *   include_self  0: sklearn's kneighbors_graph()/kneighbors() convention (query row dropped); 1: kept (distance 0)
 *   outputs as schic_mh_kneighbors / schic_mh_kneighbors_graph */
int schic_mh_kneighbors_exact(schic_mh_handle* h, int32_t k, int32_t include_self,
                              int32_t* idx, float* dist, int32_t* n_valid);
int schic_mh_kneighbors_exact_graph(schic_mh_handle* h, int32_t k, int32_t include_self,
                                    int64_t* indptr, int32_t* indices, float* data);

/* SURVEY 8(f)-4, the hyperopt fast path (scHicClusterMinHashHyperopt.py:119-148 re-runs the whole tool once per trial,
 * with -nh drawn from a range): keep ONE fitted handle (created with the largest hash count of the sweep) and switch
 * the number of hash functions the query side uses. h_j depends only on j, so the first n_hashes columns of the fitted
 * signatures are exactly the signatures a fit with number_of_hash_functions = n_hashes would produce; the CSR stays
 * resident and nothing is re-hashed. Afterwards signatures / collision_counts / kneighbors* behave as if the handle
 * had been created with n_hashes (minimal_blocks_in_common etc. unchanged). The next fit (append or not) switches back
 * to the created hash count. n_hashes in [1, number_of_hash_functions as created]. */
int schic_mh_set_active_hashes(schic_mh_handle* h, int32_t n_hashes);

#ifdef __cplusplus
}
#endif
#endif /* SCHIC_MH_H_ */
