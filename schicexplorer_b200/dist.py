"""Multi-GPU host layer: one process per GPU, cells sharded across ranks.

SURVEY.md 8(e): rows (cells) are independent for K1, so each rank hashes its own shard; the one
real exchange step is an all-gather of the signature matrix (uint32 [n/G, H] per rank) with NCCL
over NVLink (``torch.distributed``), plus -- only for the exact rerank -- an all-gather of the CSR
shards so that any candidate row is local. Each rank then answers the kNN queries of its own cells
against the gathered database (K3 -> select -> K4). No reduction collective is needed.

PyTorch is plumbing here (device buffers, process group, streams); all arithmetic is in the
kernels behind the C ABI (include/schic_mh_cuda.h).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_rows(n_rows: int, world: int, align: int = 8) -> list:
    """Contiguous, near-equal row blocks: rank r owns [bounds[r], bounds[r+1]). Inner bounds are multiples of ``align``
    when the shards are large enough not to notice (the TMA-staged K3 wants block offsets on 16-byte boundaries)."""
    if align > 1 and n_rows >= 64 * align * world:
        units, tail = divmod(n_rows, align)
        base, rem = divmod(units, world)
        bounds = [0]
        for r in range(world):
            bounds.append(bounds[-1] + (base + (1 if r < rem else 0)) * align)
        bounds[-1] += tail
        return bounds
    base, rem = divmod(n_rows, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def partition_rows_by_nnz(indptr: np.ndarray, world: int) -> list:
    """Contiguous row blocks balanced by non-zeros (prefix-sum split) to tame row-length skew."""
    n = len(indptr) - 1
    total = indptr[-1] - indptr[0]
    bounds = [0]
    for r in range(1, world):
        target = indptr[0] + total * r // world
        bounds.append(int(min(max(np.searchsorted(indptr, target, side="left"), bounds[-1]), n)))
    bounds.append(n)
    return bounds


def _gather_padded(local: torch.Tensor, counts: list, group, tail_pad: int = 0) -> torch.Tensor:
    """All-gather of unequal first-dimension shards: pad to the largest, gather, compact. ``tail_pad``: extra
    (uninitialised) elements allocated behind the result; the returned tensor is the view without them."""
    world = len(counts)
    mx = max(counts)
    tail = local.shape[1:]
    total = sum(counts)
    if local.shape[0] < mx:
        pad = torch.empty((mx,) + tuple(tail), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        local = pad
    if len(set(counts)) == 1:
        out = torch.empty((world * mx + tail_pad,) + tuple(tail), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out[:world * mx], local.contiguous(), group=group)
        return out[:world * mx]
    buf = torch.empty((world, mx) + tuple(tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view((world * mx,) + tuple(tail)), local.contiguous(), group=group)
    out = torch.empty((total + tail_pad,) + tuple(tail), dtype=local.dtype, device=local.device)
    torch.cat([buf[r, :counts[r]] for r in range(world)], dim=0, out=out[:total])
    return out[:total]


def _gather_var(local: torch.Tensor, counts: list, group, tail_pad: int = 0) -> torch.Tensor:
    """All-gather of unequal first-dimension shards WITHOUT padding or compaction: every rank receives each peer's shard
    straight into its place of the result (one grouped round of NCCL send / recv). For the big arrays of the exchange
    (packed stream, ids, values), where pad + all-gather + concatenate would move the data twice."""
    world, rank = len(counts), dist.get_rank(group)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + int(c))
    tail = tuple(local.shape[1:])
    out = torch.empty((offs[-1] + tail_pad,) + tail, dtype=local.dtype, device=local.device)
    local = local.contiguous()
    assert local.shape[0] == counts[rank]
    if counts[rank]:
        out[offs[rank]:offs[rank + 1]].copy_(local)
    ops = []
    for step in range(1, world):
        dst, src = (rank + step) % world, (rank - step) % world
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, local, dst, group=group))
        if counts[src]:
            ops.append(dist.P2POp(dist.irecv, out[offs[src]:offs[src + 1]], src, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out[:offs[-1]]


class PeerExchange:
    """CSR exchange over NVLink peer memory instead of an NCCL all-gather (experimental, opt-in: see the note in
    ShardedMinHash.__init__ -- correct, but slower than NCCL where legacy CUDA IPC mappings do not get NVLink DMA).

    Every rank owns two sets (double buffering) of *global* CSR arrays (ids, values: the concatenation of all
    shards, no padding) and shares them with its peers through CUDA IPC once. Per build a rank copies its shard
    straight into all peers' arrays at its global offset -- plain device-to-device copies over NVLink / NVSwitch,
    executed by the copy engines, so they neither take SMs from K1 / K3 nor need a compaction pass -- and a
    one-element all-reduce on the same stream serves as the completion barrier (it can only finish once every rank
    has issued it behind its own copies). Buffer set i % 2 is written in build i; a rank can start build i + 2
    only after the barrier of build i + 1, i.e. after every peer has finished reading set i % 2.
    """

    def __init__(self, device, group, shard_nnz, value_dtype=torch.float32):
        from torch.multiprocessing.reductions import reduce_tensor
        self.device, self.group = device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nnz = [int(x) for x in shard_nnz]
        self.off = [0]
        for x in self.nnz:
            self.off.append(self.off[-1] + x)
        total = self.off[-1]
        self.ix = [torch.empty(total, dtype=torch.int32, device=device) for _ in range(2)]
        self.dt = [torch.empty(total, dtype=value_dtype, device=device) for _ in range(2)]
        mine = [(reduce_tensor(a), reduce_tensor(b)) for a, b in zip(self.ix, self.dt)]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.peer_ix = [[None] * self.world for _ in range(2)]
        self.peer_dt = [[None] * self.world for _ in range(2)]
        for p in range(self.world):
            for b in range(2):
                if p == self.rank:
                    self.peer_ix[b][p], self.peer_dt[b][p] = self.ix[b], self.dt[b]
                else:
                    (f1, a1), (f2, a2) = everyone[p][b]
                    self.peer_ix[b][p], self.peer_dt[b][p] = f1(*a1), f2(*a2)
        self.stream = torch.cuda.Stream(device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.turn = 0
        dist.barrier(group=group)

    def exchange(self, local_ix: torch.Tensor, local_dt: torch.Tensor):
        """Issue the copies behind everything already queued on the current stream. Returns (ids, values, event)."""
        b = self.turn
        self.turn ^= 1
        lo, n = self.off[self.rank], self.nnz[self.rank]
        assert local_ix.numel() == n
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for step in range(self.world):                   # start with the next rank: spreads the load over the links
                p = (self.rank + 1 + step) % self.world
                self.peer_ix[b][p][lo:lo + n].copy_(local_ix, non_blocking=True)
                self.peer_dt[b][p][lo:lo + n].copy_(local_dt, non_blocking=True)
            dist.all_reduce(self.flag, group=self.group)     # completion barrier
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return self.ix[b], self.dt[b], ev


class ShardedMinHash:
    """Cells [row0, row0 + n_local) of an n_total-cell data set live on this rank.

    ``api`` is the CUDA C-ABI binding (or, in the CPU gloo tests, any object with the same
    methods); tensors live on ``device``.
    """

    def __init__(self, api, n_total: int, row_bounds: list, device, group=None, n_features=None, shard_nnz=None,
                 integer_counts: bool = False, **minhash_kwargs):
        self.api, self.group, self.device = api, group, device
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_total = n_total
        self.bounds = list(row_bounds)
        self.row0, self.n_local = self.bounds[self.rank], self.bounds[self.rank + 1] - self.bounds[self.rank]
        self.kw = dict(minhash_kwargs)
        if self.kw.get("shingle"):
            raise NotImplementedError("block combine (shingle=1) is a single-GPU option: the sharded layer exchanges raw signatures")
        self.H = int(self.kw["number_of_hash_functions"])
        self.fast = bool(self.kw.get("fast", True))
        dev_index = device.index if device.type == "cuda" else -1
        self.h = api.create(device=dev_index, **self.kw)
        if device.type == "cuda":
            api.set_stream(self.h, torch.cuda.current_stream(device).cuda_stream)
        # optional knowledge of the caller that removes host synchronisations from every build: the column count
        # of the matrix (ids < n_features) and the non-zeros of every rank's shard (sizes of the CSR exchange)
        if n_features is not None and hasattr(api, "set_feature_range"):
            api.set_feature_range(self.h, int(n_features))
        self.shard_nnz = None if shard_nnz is None else [int(x) for x in shard_nnz]
        # integer_counts: the caller states that the values are small non-negative integers (Hi-C contact counts). The
        # exact rerank then exchanges ONE packed 32-bit word per non-zero ((id mod 2^16) << 16 | count, produced by the
        # owning rank's preparation pass) instead of the id and value arrays -- half the NVLink volume. The statement is
        # verified on the device (flags of the preparation pass): if it does not hold, the next synchronising call
        # raises instead of returning wrong distances.
        self.integer_counts = bool(integer_counts)
        # The CSR exchange gets its own communicator: collectives of one NCCL communicator run in issue order on one
        # internal stream, so on the default group the (small, urgent) signature all-gather would queue behind the
        # 1 GB CSR all-gather that was issued before K1 and K3 could not start until it had finished.
        self._csr_group = None
        self._peer = None
        # rerank preparation (row norms, window directory) on the rank that owns the rows, exchanged with the CSR:
        # needs the feature-range hint (every rank must cut the id axis into the same windows)
        self._nw = 0
        if (self.world > 1 and not self.fast and device.type == "cuda" and n_features is not None
                and hasattr(api, "rerank_windows")):
            self._nw = api.rerank_windows(self.h)
        if self.world > 1 and not self.fast and device.type == "cuda" and group is None:
            self._csr_group = dist.new_group(backend="nccl")
            # SCHIC_CSR_EXCHANGE=p2p: copies into peer memory mapped through CUDA IPC instead of the NCCL all-gather
            # (needs the shard sizes). Opt-in: on this pool's boxes a copy into legacy-IPC-mapped peer memory runs
            # at 25 GB/s (tools/p2p_probe.py; same-process P2P: 735 GB/s, NCCL all-gather: 237 GB/s busbw at 2 ranks),
            # so the NCCL path stays the default.
            import os
            if self.shard_nnz is not None and os.environ.get("SCHIC_CSR_EXCHANGE", "nccl") == "p2p":
                try:
                    self._peer = PeerExchange(device, self._csr_group, self.shard_nnz)
                except Exception as e:       # no IPC / peer access on this box: the NCCL path still works
                    import warnings
                    warnings.warn(f"peer-memory CSR exchange unavailable ({e}); using the NCCL all-gather")
                    self._peer = None
        self._keep = {}
        self._side = None
        self._csr_event = None

    def close(self):
        if self.h is not None:
            self.api.destroy(self.h)
            self.h = None

    # -- step 1: hash the local shard (device-resident CSR) ---------------------------------
    def fit_local(self, d_indptr: torch.Tensor, d_indices: torch.Tensor, d_data, indptr0=None):
        """``indptr0``: the value of ``d_indptr[0]`` if the caller knows it (0 for a stand-alone shard); the fit then
        enqueues without reading it back from the device (no host synchronisation in the build)."""
        self._keep["csr"] = (d_indptr, d_indices, d_data)
        self._prep_local(d_indptr, d_indices, d_data)
        self._start_csr_gather()
        nnz = int(d_indices.numel())
        extra = {} if indptr0 is None else {"indptr0": int(indptr0)}
        self.api.fit_csr_device(self.h, self.n_local, nnz, d_indptr.data_ptr(), d_indices.data_ptr(),
                                None if d_data is None else d_data.data_ptr(), False, **extra)

    def _prep_local(self, d_indptr, d_indices, d_data):
        """Rerank preparation of this rank's rows (norms, window directory, flags, packed stream), exchanged with the CSR."""
        self._keep.pop("prep_local", None)
        if self._nw <= 0 or d_data is None:
            return
        norms = torch.empty(self.n_local, dtype=torch.float64, device=self.device)
        wdir = torch.empty((self.n_local, self._nw + 1), dtype=torch.int32, device=self.device)
        flags = torch.empty(4, dtype=torch.int32, device=self.device)
        packed = torch.empty(d_indices.numel(), dtype=torch.int32, device=self.device) if self.integer_counts else None
        self.api.rerank_prep_rows(self.h, self.n_local, d_indptr.data_ptr(), d_indices.data_ptr(), d_data.data_ptr(),
                                  self._nw, norms.data_ptr(), wdir.data_ptr(), flags.data_ptr(),
                                  None if packed is None else packed.data_ptr())
        self._keep["prep_local"] = (norms, wdir, flags, packed)

    def fit_local_host(self, indptr: np.ndarray, indices: np.ndarray, data, async_upload: bool = False):
        """Same, from HOST buffers (pinned for full PCIe speed): the H2D copy is part of the call and
        is overlapped with K1 inside the library. ``async_upload``: do not wait for the value array to
        land (only the exact rerank reads it) -- the caller keeps the buffers alive until the step's
        results have been read back."""
        if async_upload:
            self._keep["host_csr"] = (indptr, indices, data)
            self.api.fit_csr(self.h, indptr, indices, data, False, async_upload=True)
        else:
            self.api.fit_csr(self.h, indptr, indices, data, False)
        p_ip, p_ix, p_dt, nnz = self.api.csr_device(self.h)
        d_ip = _wrap_device(p_ip, (self.n_local + 1,), "<i8", torch.int64, self.device)
        d_ix = _wrap_device(p_ix, (nnz,), "<i4", torch.int32, self.device)
        d_dt = None if not p_dt else _wrap_device(p_dt, (nnz,), "<f4", torch.float32, self.device)
        self._keep["csr"] = (d_ip, d_ix, d_dt)
        if async_upload and self.world > 1 and not self.fast:
            self.api.synchronize(self.h)       # the CSR exchange reads the value array: it must have landed
        self._prep_local(d_ip, d_ix, d_dt)
        self._start_csr_gather()

    # -- step 2: the exchange ------------------------------------------------------------------
    def _gather_csr(self):
        """All-gather of the CSR shards (only the exact rerank reads other ranks' rows)."""
        counts = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        d_indptr, d_indices, d_data = self._keep["csr"]
        group = self._csr_group if self._csr_group is not None else self.group
        if self.shard_nnz is not None:
            nnzs = self.shard_nnz
            assert nnzs[self.rank] == d_indices.numel(), "shard_nnz does not match the local shard"
        else:
            nnz_local = torch.tensor([d_indices.numel()], dtype=torch.int64, device=self.device)
            nnz_all = torch.empty(self.world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(nnz_all, nnz_local, group=group)
            nnzs = [int(x) for x in nnz_all.tolist()]
        ev = None
        gather_big = _gather_var if self.device.type == "cuda" else _gather_padded
        packed_only = "prep_local" in self._keep and self._keep["prep_local"][3] is not None
        if packed_only:
            ix_all = dt_all = None            # the packed stream (gathered below) is all the rerank reads
        elif self._peer is not None:
            ix_all, dt_all, ev = self._peer.exchange(d_indices, d_data)
        else:
            ix_all = gather_big(d_indices, nnzs, group)
            dt_all = gather_big(d_data, nnzs, group)
        lens = (d_indptr[1:] - d_indptr[:-1]).contiguous()
        self._keep.pop("prep_all", None)
        if "prep_local" in self._keep:
            norms, wdir, flags, packed = self._keep["prep_local"]
            packed_all = None
            if packed is not None:
                # (16 readable bytes behind the stream: the counts kernel reads slices as aligned 16-byte words)
                packed_all = gather_big(packed, nnzs, group, tail_pad=4)
            # the small per-row arrays travel in ONE collective: [lens | norms | directory | flags] per rank, padded to the
            # largest shard (row lengths as int64, norms as float64, both viewed as bytes)
            mx, nw1 = max(counts), self._nw + 1
            seg = [(x + 15) // 16 * 16 for x in (mx * 8, mx * 8, mx * nw1 * 4, 16)]      # 16-byte aligned segments
            buf = torch.zeros(sum(seg), dtype=torch.uint8, device=self.device)
            nl = self.n_local
            buf[:nl * 8].copy_(lens.view(torch.uint8))
            buf[seg[0]:seg[0] + nl * 8].copy_(norms.view(torch.uint8))
            buf[seg[0] + seg[1]:seg[0] + seg[1] + nl * nw1 * 4].copy_(wdir.reshape(-1).view(torch.uint8))
            buf[seg[0] + seg[1] + seg[2]:].copy_(flags.view(torch.uint8))
            allb = torch.empty((self.world, sum(seg)), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allb.view(-1), buf, group=group)
            o1, o2, o3 = seg[0], seg[0] + seg[1], seg[0] + seg[1] + seg[2]
            lens_all = torch.cat([allb[r, :counts[r] * 8].view(torch.int64) for r in range(self.world)])
            norms_all = torch.cat([allb[r, o1:o1 + counts[r] * 8].view(torch.float64) for r in range(self.world)])
            dir_all = torch.cat([allb[r, o2:o2 + counts[r] * nw1 * 4].view(torch.int32).view(counts[r], nw1)
                                 for r in range(self.world)])
            flags_all = allb[:, o3:].contiguous().view(torch.int32).view(self.world, 4).sum(dim=0).to(torch.int32)   # 0/1 flags and one count: a sum combines them
            self._keep["prep_all"] = (norms_all, dir_all, flags_all, packed_all, sum(nnzs))
        else:
            lens_all = _gather_padded(lens, counts, group)
        # row lengths -> global indptr
        ip_all = torch.zeros(self.n_total + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(lens_all, 0, out=ip_all[1:])
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        return ip_all, ix_all, dt_all

    def _start_csr_gather(self):
        """The CSR exchange does not depend on the signatures: on a GPU it is issued on a side stream
        *before* K1 so that the NVLink transfer (1.15 GB at S10) runs underneath the hashing."""
        self._keep.pop("csr_all", None)
        self._csr_event = None
        if self.fast or self.world == 1 or self.device.type != "cuda":
            return
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)
        self._side.wait_stream(main)           # the previous build no longer reads the old gathered CSR
        with torch.cuda.stream(self._side):
            self._keep["csr_all"] = self._gather_csr()
            self._csr_event = torch.cuda.Event()
            self._csr_event.record(self._side)

    def gather(self):
        """All-gather signatures (and the CSR shards when the exact rerank needs them) and hand
        the gathered database back to the kernels."""
        counts = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        ptr, n = self.api.signatures_device(self.h)
        assert n == self.n_local
        sig_local = _wrap_device(ptr, (self.n_local, self.H), "<i4", torch.int32, self.device)
        if self.world == 1:
            sig_all = sig_local
        else:
            sig_all = _gather_padded(sig_local, counts, self.group)
        self._keep["sig_all"] = sig_all
        ip_all = ix_all = dt_all = None
        pending = None
        if not self.fast:
            if self.world == 1:
                ip_all, ix_all, dt_all = self._keep["csr"]
            elif "csr_all" in self._keep:
                # the exchange may still be running: the library waits for its event right before the rerank,
                # so K3 and the select (which only read signatures) overlap it
                ip_all, ix_all, dt_all = self._keep["csr_all"]
                pending = self._csr_event
            else:
                ip_all, ix_all, dt_all = self._gather_csr()
            self._keep["csr_all"] = (ip_all, ix_all, dt_all)
        self.api.set_database_device(self.h, self.n_total, self.row0, sig_all.data_ptr(),
                                     None if ip_all is None else ip_all.data_ptr(),
                                     None if ix_all is None else ix_all.data_ptr(),
                                     None if dt_all is None else dt_all.data_ptr())
        if pending is not None:
            self.api.set_database_ready_event(self.h, pending.cuda_event)
        if not self.fast and self.world > 1 and "prep_all" in self._keep:
            norms_all, dir_all, flags_all, packed_all, nnz_total = self._keep["prep_all"]
            if packed_all is None:
                self.api.set_database_prep_device(self.h, norms_all.data_ptr(), dir_all.data_ptr(), self._nw)
            else:
                self.api.set_database_prep_device(self.h, norms_all.data_ptr(), dir_all.data_ptr(), self._nw,
                                                  packed_all.data_ptr(), flags_all.data_ptr(), nnz_total)

    # -- step 2b: collision counts of a sharded run, every (rank x rank) block computed once -------------------------
    def _symmetric_counts_possible(self) -> bool:
        import os
        mode = os.environ.get("SCHIC_K3_SHARDING", "auto")
        if not (self.world > 1 and self.device.type == "cuda" and hasattr(self.api, "collision_block_device")
                and self.H <= 4094 and self.n_total <= 57344 and mode in ("auto", "symmetric")):
            return False
        # one launch per peer: pays when a rank's block row holds several waves of 128 x 128 tiles per launch (a wave =
        # 2 CTAs x 148 SMs). Measured on S10: 2 GPUs 2.43 -> 1.53 ms, 8 GPUs (1 250 rows per rank) 1.02 -> 1.9 ms.
        return mode == "symmetric" or self.n_local * self.n_total / (128.0 * 128.0 * 296.0) > 1.3 * self.world

    def _symmetric_counts(self):
        """The count matrix is symmetric. With rows sharded, rank r needs block row r of it; computing that row alone
        costs n^2 H / G compares per rank although every off-diagonal block (r, s) is the transpose of (s, r). Here rank r
        computes its diagonal block with the symmetric kernel and HALF of every off-diagonal block of its rows -- for
        s > r the columns of the first half of shard s, for s < r the rows of the second half of its own shard -- and
        the kernel stores each such piece a second time, transposed, into a send buffer: exactly the piece rank s is
        not computing. One point-to-point exchange (NCCL send/recv, ~(n/G)^2 bytes per peer) completes every rank's
        block row: n^2 H / (2 G) compares per rank. Returns the [n_local, ld] uint16 matrix (as int16 tensor)."""
        G, r, b = self.world, self.rank, self.bounds
        nl = self.n_local
        ld = (self.n_total + 7) // 8 * 8
        counts = self._keep.get("counts")
        if counts is None or counts.shape != (nl, ld):
            counts = torch.empty((nl, ld), dtype=torch.int16, device=self.device)
            self._keep["counts"] = counts
        base, esz = counts.data_ptr(), 2

        def half(s):                       # shard s = [lo, mid) + [mid, hi), mid on a multiple of 8 when it can be
            lo, hi = b[s], b[s + 1]
            mid = lo + (hi - lo + 1) // 2
            if hi - lo >= 64:
                mid = min(hi, (mid + 7) // 8 * 8)
            return lo, mid, hi

        my_lo, my_mid, my_hi = half(r)
        self.api.collision_block_device(self.h, my_lo, nl, my_lo, nl, base + my_lo * esz, ld)          # diagonal block
        sends, recvs, ops = {}, {}, []
        for s in range(G):
            if s == r:
                continue
            s_lo, s_mid, s_hi = half(s)
            if r < s:      # all my rows x first half of shard s; its transpose is what s lacks
                rows0, nrows, cols0, ncols = my_lo, nl, s_lo, s_mid - s_lo
                rshape = (nl, s_hi - s_mid)                   # from s: my rows x second half of shard s
            else:          # second half of my rows x all of shard s
                rows0, nrows, cols0, ncols = my_mid, my_hi - my_mid, s_lo, s_hi - s_lo
                rshape = (my_mid - my_lo, s_hi - s_lo)        # from s: first half of my rows x shard s
            ld_t = (nrows + 7) // 8 * 8
            snd = torch.empty((ncols, ld_t), dtype=torch.int16, device=self.device)
            self.api.collision_block_device(self.h, rows0, nrows, cols0, ncols,
                                            base + ((rows0 - my_lo) * ld + cols0) * esz, ld, snd.data_ptr(), ld_t)
            # what s sends has ITS row stride: rows = rshape[0] (mine), padded to a multiple of 8
            sends[s] = snd
            recvs[s] = torch.empty((rshape[0], (rshape[1] + 7) // 8 * 8), dtype=torch.int16, device=self.device)
        # the piece s computed for me is [its columns = my rows] transposed: it arrives as [my rows, its rows (padded)]
        for s in range(G):
            if s == r:
                continue
            ops.append(dist.P2POp(dist.isend, sends[s].view(torch.uint8), s, group=self.group))     # (NCCL has no 16-bit integer type)
            ops.append(dist.P2POp(dist.irecv, recvs[s].view(torch.uint8), s, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for s in range(G):
            if s == r:
                continue
            s_lo, s_mid, s_hi = half(s)
            rv = recvs[s]
            if r < s:      # my rows x second half of shard s
                counts[:, s_mid:s_hi].copy_(rv[:, :s_hi - s_mid])
            else:          # first half of my rows x shard s
                counts[:my_mid - my_lo, s_lo:s_hi].copy_(rv[:, :s_hi - s_lo])
        self._keep["k3_sends"] = (sends, recvs)        # keep the buffers alive until the stream has used them
        return counts, ld

    # -- step 3: local queries -------------------------------------------------------------------
    def _counts_for_queries(self):
        if self._symmetric_counts_possible() and "sig_all" in self._keep:
            counts, ld = self._symmetric_counts()
            self.api.set_counts_device(self.h, counts.data_ptr(), ld)

    def kneighbors_local(self, k: int, out=None):
        if out is None:
            out = (torch.empty((self.n_local, k), dtype=torch.int32, device=self.device),
                   torch.empty((self.n_local, k), dtype=torch.float32, device=self.device),
                   torch.empty(self.n_local, dtype=torch.int32, device=self.device))
        idx, dst, nv = out
        self._counts_for_queries()
        self.api.kneighbors_device(self.h, k, -1, idx.data_ptr(), dst.data_ptr(), nv.data_ptr())
        return out


    def kneighbors_graph_local(self, k: int, out=None):
        """``kneighbors_graph(mode='distance')`` rows of this rank's cells (columns are global cell ids), CSR left on
        the device: (indptr int64 [n_local+1], indices int32, data float32; the first indptr[-1] entries are valid)."""
        if out is None:
            out = (torch.empty(self.n_local + 1, dtype=torch.int64, device=self.device),
                   torch.empty(self.n_local * k, dtype=torch.int32, device=self.device),
                   torch.empty(self.n_local * k, dtype=torch.float32, device=self.device))
        ip, ix, dt = out
        self._counts_for_queries()
        self.api.kneighbors_graph_device(self.h, k, -1, ip.data_ptr(), ix.data_ptr(), dt.data_ptr())
        return out

    def kneighbors_exact_local(self, k: int, include_self: bool = False, out=None):
        """Exact euclidean kNN (scHicCluster -drm knn) of this rank's rows against the whole gathered
        data set: brute-force K4, no candidate generation. Rows shard with no exchange beyond the CSR all-gather
        ``gather()`` already did (needs ``fast=False`` so that the CSR is there)."""
        if self.fast:
            raise ValueError("kneighbors_exact_local needs the CSR values: create the shard with fast=False")
        if out is None:
            out = (torch.empty((self.n_local, k), dtype=torch.int32, device=self.device),
                   torch.empty((self.n_local, k), dtype=torch.float32, device=self.device),
                   torch.empty(self.n_local, dtype=torch.int32, device=self.device))
        idx, dst, nv = out
        self.api.kneighbors_exact_device(self.h, k, include_self, idx.data_ptr(), dst.data_ptr(), nv.data_ptr())
        return out


def _wrap_device(ptr: int, shape, typestr: str, dtype, device) -> torch.Tensor:
    """Zero-copy torch view of a device buffer owned by the C library (uint32 is viewed as int32)."""
    n = int(np.prod(shape))
    if n == 0:
        return torch.empty(shape, dtype=dtype, device=device)
    if device.type == "cpu":   # CPU tests (gloo): the "device" buffers are host memory
        import ctypes
        buf = (ctypes.c_char * (n * np.dtype(typestr).itemsize)).from_address(ptr)
        return torch.from_numpy(np.frombuffer(buf, dtype=np.dtype(typestr), count=n)).view(shape)

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Holder(), device=device).view(shape)
