"""``scHicClusterMinHashHyperopt`` with the data resident on the GPU across trials.

The reference (``/root/reference/schicexplorer/scHicClusterMinHashHyperopt.py:119-148``) evaluates every trial by
re-running ``scHicClusterMinHash.main`` end to end: reload the scool, rebuild the CSR, hash it with that trial's ``-nh``,
build the graph, embed, cluster, then read the score back from the file ``correct_associated``
(``scHicClusterMinHash.py:470-580``). Here one trial only costs its query side:

* the cells are loaded once, one CSR per ``intraChromosomalContactsOnly`` value the search space holds;
* each CSR is fitted ONCE, at the largest hash count of the range, and stays on the device;
* a trial switches the number of hash functions with ``MinHash.set_number_of_hash_functions`` (a prefix of the fitted
  signatures is the signature matrix of the smaller count - ``schic_mh_set_active_hashes``) and asks for the graph;
* embedding (PCA / UMAP), spectral clustering and the score are the reference's host steps.

Flags as the reference (``:13-115``). ``--method random`` is a seeded uniform draw over the same space and needs nothing
beyond numpy; ``tpe`` / ``atpe`` hand the same objective to the ``hyperopt`` package when it is installed.
"""
from __future__ import annotations

import argparse
import logging

import numpy as np

from . import __version__
from .clustering import MinHashClustering
from .minhash import MinHash
from .overlap import correct_associated, read_cell_types  # noqa: F401  (re-exported: tests and callers import them from here)
from .scHicClusterMinHash import _umap_available, load_cells, parse_arguments as _tool_arguments

log = logging.getLogger(__name__)


def parse_arguments(args=None):
    parser = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter, add_help=False,
                                     description='Hyper-parameter search for scHicClusterMinHash; the matrix is hashed '
                                                 'once and stays on the GPU across trials.')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrix', '-m', required=True, help='scool matrix (also: sparse-matrix-files folder, cells npz).')
    req.add_argument('--cellColor', '-c', required=True, help='Two-column file: cell name, cell type / cycle stage.')
    req.add_argument('--outputFileName', '-o', default='hyperoptMinHash_result.txt', required=False,
                     help='Result of the optimisation (Default: %(default)s).')
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--runs', type=int, default=100, help='Number of trials.')
    opt.add_argument('--nearestNeighbors', '-k', type=int, default=1000, help='Neighbours per cell in the kNN graph.')
    opt.add_argument('--numberOfHashfunctions', '-noh', type=int, nargs=3, default=(1000, 20000, 1000),
                     help='Hash function range: start, stop, stepsize (Default: %(default)s).')
    opt.add_argument('--numberOfClusters', '-noc', type=int, nargs=2, default=(6, 15),
                     help='Cluster count range (Default: %(default)s).')
    opt.add_argument('--numberPCADimensions', '-nop', type=int, nargs=3, default=(30, 60, 1),
                     help='PCA dimension range: start, stop, stepsize (Default: %(default)s).')
    opt.add_argument('--umap_numberOfNeighbors', '-unon', type=int, nargs=3, default=(30, 60, 1),
                     help='umap n_neighbors range: start, stop, stepsize (Default: %(default)s).')
    opt.add_argument('--umap_n_components', '-unoc', type=int, nargs=3, default=(2, 10, 1),
                     help='umap n_components range: start, stop, stepsize (Default: %(default)s).')
    opt.add_argument('--umap_min_dist', '-umin', type=float, nargs=2, default=(0.0, 0.5),
                     help='umap min_dist range: start, stop (Default: %(default)s).')
    req.add_argument('--method', '-me', choices=['random', 'tpe', 'atpe'], default='random', required=True,
                     help='Search method.')
    opt.add_argument('--threads', '-t', required=False, default=16, type=int, help='Host threads (Default: %(default)s).')
    opt.add_argument('--help', '-h', action='help', help='Show this help message and exit.')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    return parser


# the score of scHicClusterMinHash.py:470-580 lives in overlap.py (shared with the tool's -cct output)


# ---- resident state --------------------------------------------------------------------------------------------------
class ResidentCells:
    """The cells of one scool, fitted once per ``intra_only`` value at the largest hash count and kept on the device."""

    def __init__(self, matrix: str, max_hashes: int, n_neighbors: int, threads: int = 8):
        from . import loader
        self._loader = loader
        self.cells, self.chroms = load_cells(matrix, threads)
        self.names = [c.name for c in self.cells]
        self.max_hashes = int(max_hashes)
        self.n_neighbors = min(int(n_neighbors), len(self.cells))
        self.threads = threads
        self._fitted = {}
        self.fits = 0            # how often a CSR was hashed (the point of this class: at most twice per search)

    def minhash(self, intra_only: bool) -> MinHash:
        mh = self._fitted.get(bool(intra_only))
        if mh is None:
            X, _ = self._loader.cells_to_csr(self.cells, self.chroms, intra_only=bool(intra_only))
            # kwargs of the tool's default branch (scHicClusterMinHash.py:402-404)
            mh = MinHash(n_neighbors=self.n_neighbors, number_of_hash_functions=self.max_hashes,
                         number_of_cores=self.threads, shingle_size=5, fast=True, maxFeatures=int(max(X.getnnz(1))),
                         absolute_numbers=False, max_bin_size=100000, minimal_blocks_in_common=100, excess_factor=1,
                         prune_inverse_index=False)
            mh.fit(X)
            self.fits += 1
            self._fitted[bool(intra_only)] = mh
        return mh

    def graph(self, intra_only: bool, n_hashes: int):
        mh = self.minhash(intra_only).set_number_of_hash_functions(n_hashes)
        g = mh.kneighbors_graph(mode='distance')
        g.data[~np.isfinite(g.data)] = 0
        return g

    def close(self):
        for mh in self._fitted.values():
            mh.close()
        self._fitted = {}


def trial_umap_dict(params: dict) -> dict:
    """The umap arguments of one trial. The reference objective runs the tool itself, i.e. with ALL of the tool's
    ``--umap_*`` defaults (notably ``--umap_metric canberra``, scHicClusterMinHash.py:165-168, consumed at :374-382) and
    only the three searched values overridden (scHicClusterMinHashHyperopt.py:132-140); ``umap_random`` is 42 (:316)."""
    defaults = vars(_tool_arguments().parse_args(['-m', 'unused', '-o', 'unused']))
    umap_dict = {k: v for k, v in defaults.items() if 'umap_' in k}
    umap_dict.update({'umap_n_neighbors': params['umap_n_neighbors'], 'umap_n_components': params['umap_n_components'],
                      'umap_min_dist': params['umap_min_dist'], 'umap_random': 42})
    return umap_dict


def objective(params: dict, resident: ResidentCells, cell_types) -> float:
    """One trial (reference objective, :119-148): error score = 1 - correct_associated."""
    if params['dimensionsPCA'] <= params['umap_n_components']:
        return 1
    from sklearn.cluster import SpectralClustering
    graph = resident.graph(params['intraChromosomalContactsOnly'], params['numberOfHashFunctions'])
    use_umap = _umap_available()
    umap_dict = trial_umap_dict(params) if use_umap else {}
    embedding = MinHashClustering._post_process(graph, not params['noPCA'], params['dimensionsPCA'], use_umap, umap_dict)
    cluster = SpectralClustering(n_clusters=params['numberOfClusters'], affinity='nearest_neighbors',
                                 n_jobs=params['threads'], random_state=0)
    labels = cluster.fit_predict(embedding)
    error_score = 1 - correct_associated(labels, cell_types)
    print('Error score: {}'.format(error_score))
    return error_score


def search_space(args) -> dict:
    """The reference's space (:153-170) as plain value lists; a (lo, hi) tuple under 'umap_min_dist' is uniform."""
    return {
        'numberOfClusters': list(range(args.numberOfClusters[0], args.numberOfClusters[1], 1)),
        'clusterMethod': ['spectral'],
        'intraChromosomalContactsOnly': [True, False],
        'numberOfHashFunctions': list(range(*args.numberOfHashfunctions)),
        'noPCA': [False, True],
        'dimensionsPCA': list(range(*args.numberPCADimensions)),
        'umap_n_neighbors': list(range(*args.umap_numberOfNeighbors)),
        'umap_n_components': list(range(*args.umap_n_components)),
        'umap_min_dist': tuple(args.umap_min_dist),
    }


def _random_search(space, fixed, evaluate, runs, seed=0):
    rng = np.random.default_rng(seed)
    best, best_score, trials = None, np.inf, []
    for _ in range(runs):
        p = dict(fixed)
        for key, values in space.items():
            p[key] = float(rng.uniform(*values)) if isinstance(values, tuple) else values[int(rng.integers(len(values)))]
        score = evaluate(p)
        trials.append((p, score))
        if score < best_score:
            best, best_score = p, score
    return best, best_score, trials


def main(args=None):
    args = parse_arguments().parse_args(args)
    space = search_space(args)
    if not space['numberOfHashFunctions']:
        raise SystemExit('empty --numberOfHashfunctions range')
    resident = ResidentCells(args.matrix, max(space['numberOfHashFunctions']), args.nearestNeighbors, args.threads)
    types = read_cell_types(args.cellColor)
    cell_types = [types[name] for name in resident.names]
    fixed = {'numberOfNearestNeighbors': args.nearestNeighbors, 'matrixFile': args.matrix,
             'cell_coloring_type': args.cellColor, 'threads': args.threads}

    def evaluate(p):
        return objective(p, resident, cell_types)

    try:
        if args.method == 'random':
            best, _, _ = _random_search(space, fixed, evaluate, args.runs)
        else:
            try:
                from hyperopt import Trials, atpe, fmin, hp, space_eval, tpe
            except ImportError:
                raise SystemExit('--method {} needs the hyperopt package; use --method random'.format(args.method))
            hp_space = {k: (hp.uniform(k, *v) if isinstance(v, tuple) else hp.choice(k, v)) for k, v in space.items()}
            hp_space.update(fixed)
            algo = tpe.suggest if args.method == 'tpe' else atpe.suggest
            best = space_eval(hp_space, fmin(evaluate, hp_space, algo=algo, max_evals=args.runs, trials=Trials()))
    finally:
        resident.close()
    with open(args.outputFileName, 'w') as fh:
        fh.write("# Created by scHiCExplorer schicClusterMinHashHyperopt {}\n\n".format(__version__))
        fh.write("{}".format(best))
    return best


if __name__ == '__main__':
    main()
