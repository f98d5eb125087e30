"""GPU experiment: INT32 issue-rate micro-benchmark + K1 instruction-selection variants
(include/schic_mh_cuda.h: schic_int32_peak, schic_k1_variant_ms). Writes a JSON summary."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import _capi, synth  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    api = _capi.cuda_api()
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    peak = api.int32_peak(0, 2048)
    print("int32 peak", json.dumps(peak))
    cfg = synth.CONFIGS["s10"]
    ip, ix, dt, _ = synth.generate_csr(cfg, 0, n)
    H = 800
    d_ip = torch.from_numpy(ip).to(dev)
    d_ix = torch.from_numpy(ix.view(np.int32)).to(dev)
    d_sig = torch.empty((n, H), dtype=torch.int32, device=dev)
    nnz = int(ip[-1])
    res = {}
    ref = None
    for v in [-1, 100, 200]:
        ms = api.k1_variant_ms(v, n, nnz, d_ip.data_ptr(), d_ix.data_ptr(), H, d_sig.data_ptr(), reps=3)
        sig = d_sig.cpu().numpy()
        if ref is None:
            ref = sig.copy()
        same = bool(np.array_equal(sig, ref))
        evals = nnz * H / (ms * 1e-3)
        res[v] = dict(ms=ms, gevals_per_s=evals / 1e9, same=same)
        print(f"variant {v:2d}: {ms:8.3f} ms  {evals/1e9:8.1f} Gevals/s  same={same}")
    out = dict(peak=peak, n_cells=n, nnz=nnz, H=H, variants=res)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/k1_variants.json", "w"), indent=1)
