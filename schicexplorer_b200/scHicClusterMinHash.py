"""``scHicClusterMinHash`` command line tool -- same flags, defaults and output as the reference
(``/root/reference/schicexplorer/scHicClusterMinHash.py:33-242`` flags, ``:245-631`` main), with
the approximate-kNN hot path served by the B200 library.

Differences that are forced by the environment, not by design:
* ``--matrix`` accepts, besides ``.scool`` (through :mod:`schicexplorer_b200.scool`, a cooler-free
  reader), a ``sparse-matrix-files`` folder written by ``scHicConvertFormat`` or a cells ``.npz``.
* UMAP and the scatter plots need ``umap-learn`` / ``matplotlib`` on the host; when they are not
  installed the step is skipped with a warning (the reference would fail at import time).
* sklearn's ``KMeans`` lost ``n_jobs`` / ``precompute_distances``; the call is restated
  for current sklearn with the same ``random_state=0``.
"""
from __future__ import annotations

import argparse
import logging
import os

import numpy as np

from . import __version__
from .clustering import MinHashClustering
from .minhash import MinHash

log = logging.getLogger(__name__)

CLUSTER_METHODS = ['spectral', 'kmeans', 'agglomerative_ward', 'agglomerative_complete', 'agglomerative_average',
                   'agglomerative_single', 'birch']


def parse_arguments(args=None):
    """Returns the parser (like the reference, so documentation tooling can reuse it)."""
    parser = argparse.ArgumentParser(
        formatter_class=argparse.ArgumentDefaultsHelpFormatter, add_help=False,
        description='Clusters single-cell Hi-C matrices on a MinHash approximate k-nearest-neighbour graph '
                    '(samples x samples instead of samples x bins^2), computed on a B200 GPU.')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrix', '-m', required=True, metavar='scool scHi-C matrix',
                     help='Single-cell Hi-C matrices to cluster (scool; also: sparse-matrix-files folder, cells npz).')
    req.add_argument('--numberOfClusters', '-c', required=False, default=12, type=int, help='Number of clusters.')
    req.add_argument('--clusterMethod', '-cm', choices=CLUSTER_METHODS, default='spectral', help='Cluster algorithm.')
    req.add_argument('--outFileName', '-o', required=True, default='clusters.txt',
                     help='File with one "<cell name> <cluster id>" line per cell.')

    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--euclideanModeMinHash', '-em', action='store_false',
                     help='Rerank the MinHash candidates with the exact euclidean distance.')
    opt.add_argument('--intraChromosomalContactsOnly', '-ic', action='store_true',
                     help='Use only intra-chromosomal contacts.')
    opt.add_argument('--saveIntermediateRawMatrix', '-sirm', required=False,
                     help='Save the (cells x bins^2) matrix as npz to this path.')
    opt.add_argument('--createScatterPlot', '-csp', required=False, default='scatterPlot.eps',
                     help='Scatter plot of the first two embedding dimensions.')
    opt.add_argument('--dpi', required=False, default=300, type=int, help='Plot resolution.')
    opt.add_argument('--numberOfHashFunctions', '-nh', required=False, default=4000, type=int,
                     help='Number of hash functions of the MinHash signatures.')
    opt.add_argument('--numberOfNearestNeighbors', '-k', required=False, default=None, type=int,
                     help='Neighbours per cell in the kNN graph (default: all cells).')
    opt.add_argument('--shareOfMatrixToBeTransferred', '-s', required=False, default=0.25, type=float,
                     help='Share of the matrix handed to the hashing back end per chunk.')
    opt.add_argument('--saveMemory', '-sm', action='store_true',
                     help='Load and hash the cells chunk by chunk.')
    opt.add_argument('--cell_coloring_type', '-cct', required=False,
                     help='Two-column file: cell name, cell type. Writes the cluster/type overlap (matches.txt, or the '
                          'table of -lt) and the score file correct_associated; colours a second scatter plot.')
    opt.add_argument('--cell_coloring_batch', '-ccb', required=False,
                     help='Two-column file: cell name, batch. Only colours a scatter plot (needs matplotlib).')
    opt.add_argument('--noPCA', action='store_true', help='Skip the PCA of the kNN graph.')
    opt.add_argument('--noUMAP', action='store_true', help='Skip the UMAP embedding.')
    opt.add_argument('--dimensionsPCA', '-dim_pca', default=100, type=int, help='PCA dimensions kept.')
    opt.add_argument('--colorMap', default='glasbey_dark', help='Colour map of the plots.')
    opt.add_argument('--fontsize', type=float, default=15, help='Plot font size.')
    opt.add_argument('--distance', '-d', type=float, default=None,
                     help='Use only contacts closer than this genomic distance (bp).')
    opt.add_argument('--figuresize', type=float, nargs=2, default=(15, 6), metavar=('x-size', 'y-size'),
                     help='Figure size.')
    opt.add_argument('--chromosomes', nargs='+', help='Restrict to these chromosomes.')
    opt.add_argument('--absoluteValues', '-av', help='Accepted for compatibility (parsed but unused, as in the reference).')
    opt.add_argument('--latexTable', '-lt', help='With -cct: write the cluster/type overlap as a LaTeX table to this file '
                                                 '(instead of matches.txt).')
    opt.add_argument('--runInHyperoptMode', action='store_true', help='Accepted for compatibility (unused).')
    opt.add_argument('--threads', '-t', required=False, default=8, type=int, help='Host threads.')
    opt.add_argument('--help', '-h', action='help', help='show this help message and exit')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))

    um = parser.add_argument_group('Optional umap arguments', 'Passed through to umap.UMAP.')
    um.add_argument('--umap_n_neighbors', type=int, default=30)
    um.add_argument('--umap_n_components', type=int, default=2)
    um.add_argument('--umap_metric', type=str, default='canberra')
    um.add_argument('--umap_n_epochs', type=int, default=None)
    um.add_argument('--umap_learning_rate', type=float, default=1.0)
    um.add_argument('--umap_init', type=str, default='spectral')
    um.add_argument('--umap_min_dist', type=float, default=0.3)
    um.add_argument('--umap_spread', type=float, default=1.0)
    um.add_argument('--umap_set_op_mix_ratio', type=float, default=1.0)
    um.add_argument('--umap_local_connectivity', type=float, default=1.0)
    um.add_argument('--umap_repulsion_strength', type=float, default=1.0)
    um.add_argument('--umap_negative_sample_rate', type=int, default=5)
    um.add_argument('--umap_transform_queue_size', type=float, default=4.0)
    um.add_argument('--umap_a', type=float, default=None)
    um.add_argument('--umap_b', type=float, default=None)
    um.add_argument('--umap_angular_rp_forest', action='store_true')
    um.add_argument('--umap_target_n_neighbors', type=int, default=-1)
    um.add_argument('--umap_target_metric', type=str, default='categorical')
    um.add_argument('--umap_target_weight', type=float, default=0.5)
    um.add_argument('--umap_force_approximation_algorithm', action='store_true')
    um.add_argument('--umap_verbose', action='store_true')
    um.add_argument('--umap_unique', action='store_true')
    return parser


# ---- input ---------------------------------------------------------------------------------------
def load_cells(path: str, threads: int = 1):
    """Per-cell pixel tables + chromosome table from any supported container (``threads``: concurrent scool reads)."""
    from . import loader
    if os.path.isdir(path):
        return loader.read_sparse_matrix_files(path)
    if path.endswith('.npz'):
        return loader.load_cells_npz(path)
    from . import scool
    return scool.read_scool(path, threads=threads)


def make_cluster_object(method: str, n_clusters: int, threads: int):
    from sklearn.cluster import AgglomerativeClustering, Birch, KMeans, SpectralClustering
    if method == 'spectral':
        return SpectralClustering(n_clusters=n_clusters, affinity='nearest_neighbors', n_jobs=threads, random_state=0)
    if method == 'kmeans':
        return KMeans(n_clusters=n_clusters, random_state=0, n_init=10)
    if method.startswith('agglomerative'):
        return AgglomerativeClustering(n_clusters=n_clusters, linkage=method.split('_', 1)[1])
    if method == 'birch':
        return Birch(n_clusters=n_clusters)
    raise ValueError('No valid cluster method given: {}'.format(method))


def _umap_available() -> bool:
    try:
        import umap  # noqa: F401
        return True
    except ImportError:
        return False


def main(args=None):
    args = parse_arguments().parse_args(args)
    from . import loader
    cells, chroms = load_cells(args.matrix, args.threads)
    if args.numberOfNearestNeighbors is None:
        args.numberOfNearestNeighbors = len(cells)           # the kNN graph is dense by default

    cluster_object = make_cluster_object(args.clusterMethod, args.numberOfClusters, args.threads)
    use_umap = not args.noUMAP
    if use_umap and not _umap_available():
        log.warning('umap-learn is not installed: continuing without the UMAP embedding (as with --noUMAP)')
        use_umap = False
    umap_params_dict = {}
    if use_umap:
        umap_params_dict = {k: v for k, v in vars(args).items() if 'umap_' in k}
        umap_params_dict['umap_random'] = 42

    kw = dict(chromosomes=args.chromosomes, intra_only=args.intraChromosomalContactsOnly)
    names = [c.name for c in cells]
    if args.saveMemory:
        # chunked load + fit / partial_fit, then the post-processing spelled out in-tree
        max_nnz = max((len(c.bin1) for c in cells), default=0)
        per_run = max(1, int(len(cells) * args.shareOfMatrixToBeTransferred))
        mh = MinHash(n_neighbors=args.numberOfNearestNeighbors, number_of_hash_functions=args.numberOfHashFunctions,
                     number_of_cores=args.threads, shingle_size=0, fast=args.euclideanModeMinHash,
                     maxFeatures=int(max_nnz), absolute_numbers=False)
        for j, i in enumerate(range(0, len(cells), per_run)):
            X, _ = loader.cells_to_csr(cells[i:i + per_run], chroms, **kw)
            mh.fit(X) if j == 0 else mh.partial_fit(X=X)
        graph = mh.kneighbors_graph(mode='distance')
        graph.data[~np.isfinite(graph.data)] = 0
        mhc = MinHashClustering(minHashObject=mh, clusteringObject=cluster_object)
        graph = mhc._post_process(graph, not args.noPCA, args.dimensionsPCA, use_umap, umap_params_dict)
        mhc._fit_clustering(graph)
        mhc._precomputed_graph = graph
    else:
        X, _ = loader.cells_to_csr(cells, chroms, distance=args.distance, **kw)
        if args.saveIntermediateRawMatrix:
            from scipy.sparse import save_npz
            save_npz(args.saveIntermediateRawMatrix, X)
        mh = MinHash(n_neighbors=args.numberOfNearestNeighbors, number_of_hash_functions=args.numberOfHashFunctions,
                     number_of_cores=args.threads, shingle_size=5, fast=args.euclideanModeMinHash,
                     maxFeatures=int(max(X.getnnz(1))), absolute_numbers=False, max_bin_size=100000,
                     minimal_blocks_in_common=100, excess_factor=1, prune_inverse_index=False)
        mhc = MinHashClustering(minHashObject=mh, clusteringObject=cluster_object)
        mhc.fit(X=X, pSaveMemory=args.shareOfMatrixToBeTransferred, pPca=(not args.noPCA),
                pPcaDimensions=args.dimensionsPCA, pUmap=use_umap, pUmapDict=umap_params_dict)

    if args.noPCA and not use_umap and hasattr(mhc._precomputed_graph, 'data'):
        mhc._precomputed_graph.data[~np.isfinite(mhc._precomputed_graph.data)] = 0
    labels = mhc.predict(mhc._precomputed_graph, pPca=args.noPCA, pPcaDimensions=args.dimensionsPCA)

    if args.createScatterPlot:
        try:
            _scatter_plot(args, mhc._precomputed_graph, labels)
        except ImportError:
            log.warning('matplotlib is not installed: skipping %s', args.createScatterPlot)
        if args.cell_coloring_batch:
            log.warning('-ccb/--cell_coloring_batch only colours an extra scatter plot in the reference; that plot is '
                        'not produced here (no other output depends on it)')
        if args.cell_coloring_type:
            # the non-plot outputs of -cct (reference :470-580): overlap table / matches.txt and the hyperopt score file
            from . import overlap
            score = overlap.report(names, labels, args.cell_coloring_type, latex_table=args.latexTable)
            log.info('correct_associated: %s', score)
    elif args.cell_coloring_type or args.cell_coloring_batch or args.latexTable:
        log.warning('-cct / -ccb / -lt have no effect without --createScatterPlot (as in the reference)')
    if args.latexTable and not args.cell_coloring_type:
        log.warning('-lt/--latexTable needs -cct/--cell_coloring_type: no table written')

    matrices_cluster = np.array(list(zip(names, labels)))
    np.savetxt(args.outFileName, matrices_cluster, fmt="%s")
    return labels


def _scatter_plot(args, embedding, labels):
    import matplotlib
    matplotlib.use('Agg')
    import matplotlib.pyplot as plt
    emb = np.asarray(embedding.todense()) if hasattr(embedding, 'todense') else np.asarray(embedding)
    if emb.shape[1] > 2:
        from sklearn.decomposition import PCA
        emb = PCA(n_components=2).fit_transform(emb)
    plt.figure(figsize=tuple(args.figuresize))
    for lab in sorted(set(labels.tolist())):
        m = labels == lab
        plt.scatter(emb[m, 0], emb[m, 1], s=20, alpha=0.8, label=str(lab))
    plt.legend(fontsize=args.fontsize)
    plt.tight_layout()
    plt.savefig(args.createScatterPlot, dpi=args.dpi)
    plt.close()


if __name__ == '__main__':
    main()
