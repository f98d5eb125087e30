"""CPU suite: the cooler-free .scool reader against the reference's own fixtures.

tests/golden/test_matrix.scool and test_matrix_old.scool are byte copies of the reference's test data
(schicexplorer/test/test-data/); fixture20.npz is the text dump of the same 20 cells that
scHicConvertFormat wrote (test-data/scHicConvertFormat/sparse-matrix-files), so the two must agree
pixel for pixel."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from schicexplorer_b200 import loader, scool

NEW = os.path.join(GOLDEN, "test_matrix.scool")
OLD = os.path.join(GOLDEN, "test_matrix_old.scool")


def test_scool_equals_text_dump(fixture20):
    cells_txt, chroms_txt = loader.load_cells_npz(os.path.join(GOLDEN, "fixture20.npz"))
    cells, chroms = scool.read_scool(NEW)
    assert [c.name for c in cells] == [c.name for c in cells_txt] == fixture20[1]
    assert chroms.names == chroms_txt.names and chroms.binsize == chroms_txt.binsize == 1_000_000
    assert chroms.nbins == 2744 and np.array_equal(chroms.sizes, chroms_txt.sizes)
    total = 0
    for a, b in zip(cells, cells_txt):
        assert np.array_equal(a.bin1, b.bin1) and np.array_equal(a.bin2, b.bin2)
        assert np.array_equal(a.count, b.count)
        total += len(a.bin1)
    assert total == 342382
    assert len(cells[0].bin1) == 55498          # docs/content/Example_analysis.rst:296


def test_scool_csr_equals_fixture_csr(fixture20):
    cells, chroms = scool.read_scool(NEW)
    X, names = loader.cells_to_csr(cells, chroms)
    Xf = fixture20[0]
    assert names == fixture20[1] and X.shape == Xf.shape
    assert np.array_equal(X.indptr, Xf.indptr) and np.array_equal(X.indices, Xf.indices)
    assert np.array_equal(X.data, Xf.data)


def test_cell_names_and_range():
    names = scool.list_scool_cells(NEW)
    assert len(names) == 20 and all(n.startswith("/cells/") for n in names) and names == sorted(names)
    part, _ = scool.read_scool(NEW, cell_range=(5, 9))
    assert [c.name for c in part] == names[5:9]


def test_old_root_level_layout():
    """utilities.py:37-44: the old format keeps one cooler per cell in the root group, names '/<cell>'."""
    cells, chroms = scool.read_scool(OLD)
    assert len(cells) == 20 and chroms.nbins == 2744
    assert all(c.name.startswith("/") and not c.name.startswith("/cells/") for c in cells)
    for c in cells:
        assert len(c.bin1) == len(c.bin2) == len(c.count) > 0
        assert np.all(c.bin1 <= c.bin2) and c.bin2.max() < chroms.nbins


def test_rejects_non_hdf5(tmp_path):
    p = tmp_path / "x.scool"
    p.write_bytes(b"not hdf5 at all")
    with pytest.raises(scool.ScoolFormatError):
        scool.read_scool(str(p))
