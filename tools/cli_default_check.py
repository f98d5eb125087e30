import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from schicexplorer_b200 import _capi, synth
from oracle import oracle_api
from conftest import Fitted
cuda, oracle = _capi.cuda_api(), oracle_api()
cfg = synth.SynthConfig(**{**synth.CONFIGS["s10"].__dict__, "n_cells": 500})
ip, ix, dt, _ = synth.generate_csr(cfg, 0, 500, workers=16)
X = synth.to_scipy(cfg, ip, ix, dt)
n = 500
kw = dict(number_of_hash_functions=4000, minimal_blocks_in_common=100, max_bin_size=100000, excess_factor=1)
os.environ["SCHIC_K1_MODE"] = "table"
t = time.time()
with Fitted(cuda, X, fast=True, n_neighbors=n, **kw) as g:
    gi, gc, gd = cuda.kneighbors_graph(g.h, n); sg = g.sig()
print("gpu", time.time() - t)
t = time.time()
with Fitted(oracle, X, fast=True, n_neighbors=n, **kw) as o:
    oi, oc, od = oracle.kneighbors_graph(o.h, n); so = o.sig()
print("oracle", time.time() - t)
print("H=4000 k=n: signatures equal", np.array_equal(sg, so), "graph equal", np.array_equal(gi, oi) and np.array_equal(gc, oc) and np.array_equal(gd, od), "nnz", gi[-1])
