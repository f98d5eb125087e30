"""Cluster / cell-type overlap statistics of ``scHicClusterMinHash`` (the non-plot outputs of ``-cct`` / ``-lt``).

Restates ``/root/reference/schicexplorer/scHicClusterMinHash.py:258-272`` (the two-column type file), ``:470-536``
(``matches.txt`` or, with ``--latexTable``, the LaTeX table) and ``:538-580`` (the score file ``correct_associated`` that
``scHicClusterMinHashHyperopt`` reads back, ``scHicClusterMinHashHyperopt.py:142-148``). Host-side bookkeeping only.
"""
from __future__ import annotations

import numpy as np


def read_cell_types(path: str) -> dict:
    """cell name -> cell type from a two-column file (tab or four blanks), as scHicClusterMinHash.py:258-272."""
    out = {}
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            try:
                name, kind = line.split('\t')
            except ValueError:
                name, kind = line.split('    ')
            out[name] = kind
    return out


def _kinds(cell_types):
    return list(dict.fromkeys(np.asarray(cell_types).tolist()))     # first-appearance order, as the reference numbers them


def correct_associated(labels, cell_types, percentage_threshold: float = 0.8) -> float:
    """Share of each cell type that sits in clusters dominated (>= 80 %) by that type, averaged over the types
    (scHicClusterMinHash.py:538-572); 1 - this is the hyperopt error score."""
    labels = np.asarray(labels)
    cell_types = np.asarray(cell_types)
    kinds = _kinds(cell_types)
    detected = dict.fromkeys(kinds, 0.0)
    for cluster in set(labels.tolist()):
        in_cluster = labels == cluster
        for kind in kinds:
            matches = int(np.sum(in_cluster & (cell_types == kind)))
            if matches / np.sum(in_cluster) >= percentage_threshold:
                detected[kind] += matches
    return float(np.mean([detected[kind] / np.sum(cell_types == kind) for kind in kinds]))


def write_matches(path: str, labels, cell_types) -> None:
    """One line per (computed cluster, cell type) pair, a blank line between clusters (:519-536)."""
    labels = np.asarray(labels)
    cell_types = np.asarray(cell_types)
    kinds = _kinds(cell_types)
    with open(path, 'w') as fh:
        for cluster in set(labels.tolist()):
            in_cluster = labels == cluster
            size = int(np.sum(in_cluster))
            for kind in kinds:
                of_kind = cell_types == kind
                matches = int(np.sum(in_cluster & of_kind))
                fh.write('Computed cluster {} (size: {}) matching with cell type {} (size: {}) {} times. '
                         'Rate (matches/computed_clusters): {}%\n'.format(cluster, size, kind, int(np.sum(of_kind)),
                                                                          matches, matches / size))
            fh.write('\n')


def write_latex_table(path: str, labels, cell_types) -> None:
    """The ``--latexTable`` variant (:475-517): cluster x type counts and shares, then per threshold (70/80/90 %) how many
    cells of each type sit in clusters that the type dominates."""
    labels = np.asarray(labels)
    cell_types = np.asarray(cell_types)
    kinds = _kinds(cell_types)
    thresholds = [0.7, 0.8, 0.9]
    found = {t: dict.fromkeys(kinds, 0) for t in thresholds}
    head = '\\begin{table}[!htb]\n\\footnotesize\n\\begin{tabular}{|l' + '|c' * len(kinds) + '|}\n'
    body = '\\hline Cluster ' + ''.join('& {} ({} cells)'.format(kind, int(np.sum(cell_types == kind))) for kind in kinds)
    body += '\\\\\n'
    for cluster in set(labels.tolist()):
        in_cluster = labels == cluster
        size = int(np.sum(in_cluster))
        body += '\\hline Cluster {} ({} cells)'.format(cluster, size)
        for kind in kinds:
            matches = int(np.sum(in_cluster & (cell_types == kind)))
            body += '& {} {} / {:.2f} \\% '.format(matches, 'cell' if matches == 1 else 'cells', matches / size * 100)
            for t in thresholds:
                if matches / size >= t:
                    found[t][kind] += matches
        body += '\\\\\n'
    body += '\\hline ' + '&' * len(kinds) + '\\\\\n'
    for t in thresholds:
        body += '\\hline Correct identified $>{}\\%$'.format(int(t * 100))
        for kind in kinds:
            total = int(np.sum(cell_types == kind))
            body += '& {} / {} ({:.2f} \\%)'.format(found[t][kind], total, found[t][kind] / total * 100)
        body += '\\\\\n'
    body += '\\hline \n\\end{tabular}\n\\caption{}\n\\end{table}'
    with open(path, 'w') as fh:
        fh.write(head)
        fh.write(body)


def report(names, labels, type_file: str, latex_table=None, matches_path='matches.txt',
           score_path='correct_associated') -> float:
    """Everything ``-cct`` writes besides the plot: the table (``-lt``) or ``matches.txt``, and ``correct_associated``."""
    types = read_cell_types(type_file)
    missing = [n for n in names if n not in types]
    if missing:
        raise KeyError('cells without a type in {}: {} ...'.format(type_file, missing[:3]))
    cell_types = [types[n] for n in names]
    if latex_table:
        write_latex_table(latex_table, labels, cell_types)
    else:
        write_matches(matches_path, labels, cell_types)
    score = correct_associated(labels, cell_types)
    with open(score_path, 'w') as fh:
        fh.write(str(score))
    return score
