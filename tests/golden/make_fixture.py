"""Regenerates tests/golden/fixture20.npz from the reference's own HDF5-free test data:
/root/reference/schicexplorer/test/test-data/scHicConvertFormat/sparse-matrix-files
(the 20 cells of test_matrix.scool dumped by scHicConvertFormat.py:79-99 and asserted by
test_scHicConvertFormat.py:54-68). Run in the build container only (the GPU box has no
/root/reference); the npz is committed."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from schicexplorer_b200.loader import read_sparse_matrix_files, save_cells_npz  # noqa: E402

SRC = "/root/reference/schicexplorer/test/test-data/scHicConvertFormat/sparse-matrix-files"
if __name__ == "__main__":
    cells, chroms = read_sparse_matrix_files(SRC, binsize=1_000_000)
    out = os.path.join(os.path.dirname(__file__), "fixture20.npz")
    save_cells_npz(out, cells, chroms)
    print(out, len(cells), "cells", chroms.nbins, "bins", sum(len(c.bin1) for c in cells), "pixels")
