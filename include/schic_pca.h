/* schic_pca.h -- C ABI of the dense-graph PCA post-step.
 *
 * Replaces, for the kNN graph the MinHash path produced, the host call
 *     pca = PCA(n_components=min(neighborhood_matrix.shape) - 1)
 *     neighborhood_matrix = pca.fit_transform(neighborhood_matrix.todense())[:, :dimensionsPCA]
 * of /root/reference/schicexplorer/scHicClusterMinHash.py:358-366 (the same inside sparse_neighbors_search's
 * MinHashClustering.fit, :406) and scHicCluster.py:183-188 -- scikit-learn's full-SVD PCA of the densified n x n
 * graph, which dominates the tool's wall time for the default k = n from a few thousand cells on.
 *
 * What is computed (float64 throughout, all on the device; exported by libschic_mh.so, failing with
 * SCHIC_ERR_NO_DEVICE without a GPU): column means mu; Xc = X - mu; C = Xc^T Xc; Householder tridiagonalisation of C;
 * the n_components largest eigenvalues of the tridiagonal matrix by Sturm bisection; their eigenvectors by inverse
 * iteration with re-orthogonalisation; back-transformation; sign convention of scikit-learn (svd_flip with
 * u_based_decision=False: the entry of largest magnitude of every component is positive); out = Xc V.
 * Because only the leading columns are consumed downstream, only those eigenpairs are extracted; they are the same
 * numbers the full decomposition yields (tests: 1e-9 relative to scikit-learn on the reference's call).
 *
 * Errors: as include/schic_mh.h (int status, message via schic_mh_last_error()).
 */
#ifndef SCHIC_PCA_H_
#define SCHIC_PCA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PCA(...).fit_transform(X)[:, :n_components] of a dense row-major host matrix X [n_rows, n_cols] (float64).
 *   n_components     1 .. min(n_rows, n_cols); the caller passes min(dimensionsPCA, min(shape) - 1)
 *   device           CUDA ordinal, -1 = current
 *   out              [n_rows, n_components] float64, row major: the embedding (U * S)
 *   singular_values  [n_components] float64 or NULL (PCA.singular_values_)
 *   components       [n_components, n_cols] float64 or NULL (PCA.components_)
 *   mean             [n_cols] float64 or NULL (PCA.mean_: the column means that were subtracted) */
int schic_pca_fit_transform(const double* X, int64_t n_rows, int64_t n_cols, int32_t n_components, int32_t device,
                            double* out, double* singular_values, double* components, double* mean);

/* The same for X given as the CSR the graph entry points return (schic_mh_kneighbors_graph: indptr int64 [n_rows+1],
 * indices int32, data float32), densified on the device (missing entries 0, non-finite values 0 as the reference's
 * nan_to_num / isinf scrub, scHicClusterMinHash.py:356-357). */
int schic_pca_fit_transform_csr(int64_t n_rows, int64_t n_cols, const int64_t* indptr, const int32_t* indices,
                                const float* data, int32_t n_components, int32_t device, double* out,
                                double* singular_values, double* components, double* mean);

/* Device time of the stages of the last call on this thread, milliseconds:
 * {upload+densify+centre, covariance, tridiagonalisation, eigenvalues, eigenvectors, back-transform+projection, total}. */
int schic_pca_stage_ms(float* out, int32_t n_out /* >= 7 */);

#ifdef __cplusplus
}
#endif
#endif /* SCHIC_PCA_H_ */
