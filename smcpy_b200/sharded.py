"""
Multi-GPU resampling: particles are sharded in contiguous blocks, one process per GPU
(replaces the MPI scatter/allgather of model evaluations in
smcpy/mcmc/parallel_vector_mcmc.py:36-47 by sharding the whole particle loop).

resample_sharded (default): owner-computes over NVLink peer memory -- local CDF scan, rank totals through
the peer-memory exchange kernel, then every rank searches its own CDF for the contiguous range of global
output slots whose sorted uniform falls into its segment and writes the selected rows straight into the
buffer of the rank that holds the slot.  No routing of requests, no NCCL payload, no host sync.
_resample_sharded_routed (fallback without symmetric memory, and the round-1 form): uniforms of a rank's own
slots are routed to the rank whose CDF segment holds them (all-to-all), answered, and the rows travel back.
The uniforms of every built-in resampler are sorted (stratified / systematic by construction, multinomial by
exponential spacings), so both forms deal in contiguous ranges.
"""

import os
import warnings

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine
from ._lib import call, ptr, stream_ptr


class _Phases:
    """Sub-phase CUDA-event timers (engine.PROFILE set: tools/c5_run.py); free otherwise."""

    def __init__(self):
        self.cur = None

    def __call__(self, name):
        if engine.PROFILE is None:
            return
        if self.cur is not None:
            self.cur.__exit__()
        self.cur = engine._timed(name) if name else None
        if self.cur is not None:
            self.cur.__enter__()


def segment_offsets(totals):
    """O_0 = 0, O_r = W_0 + ... + W_{r-1} summed in rank order (identical on every rank)."""
    off = torch.zeros(totals.shape[0] + 1, dtype=torch.float64, device=totals.device)
    off[1:] = torch.cumsum(totals, dim=0)
    return off


def route_counts(dest, world):
    """Rows this rank sends to each destination (dest is a sorted int64 tensor)."""
    return torch.bincount(dest, minlength=world)[:world]


def route_counts_sorted(u, bounds):
    """The same counts for SORTED uniforms without touching every slot: destinations are
    non-decreasing, so the counts are the differences of the positions of the segment boundaries
    inside u (slot goes to rank r iff bounds[r-1] <= u < bounds[r])."""
    cum = torch.searchsorted(u, bounds, right=False)
    return torch.diff(cum, prepend=torch.zeros(1, dtype=cum.dtype, device=u.device))


def exchange_counts(send_counts):
    """all-to-all of the per-destination counts -> per-source counts (works on any backend)."""
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts)
    return recv


def sorted_uniforms(st, seed, stream_id):
    """st.u <- this shard's block of the sorted `standard` uniforms (smcb_resample_exponentials ->
    smcb_cdf_scan -> rank offsets -> smcb_resample_spacings_finish); st.w_alt is scratch."""
    rank, world = engine.dist_info()
    s = stream_ptr()
    n = st.n
    t = st.w_alt
    call("smcb_resample_exponentials", st.id_offset, n, st.n_global, seed, stream_id, ptr(st.u), s)
    call("smcb_cdf_scan", ptr(st.u), ptr(t), n, _lib.SCAN_LOOKBACK, ptr(st.ws), s)
    if world > 1:
        off = segment_offsets(engine._all_gather_rows(t[n - 1:n]).reshape(-1))
        lo, tot = off[rank:rank + 1], off[world:world + 1]
    else:
        lo, tot = st.zero, t[n - 1:n]          # one shard: offset 0, total = the last running sum
    call("smcb_resample_spacings_finish", ptr(t), n, ptr(lo), ptr(tot), st.n_global, seed, stream_id, ptr(st.u), s)
    return st.u


def resample_sharded(st, kernel, resample_rng, warn_threshold, defer=False):
    """Owner-computes resampling of sharded particles (reference semantics
    smcpy/smc/updater.py:135-165): no request routing, no NCCL payload, no host round trip.
      1. local look-back scan of the (globally normalised) weights, straight from the resident
         columns when the reweight left them pending                    [smcb_cdf_scan_weights]
         (+ running sums of the slot exponentials for multinomial)       [smcb_resample_exponential_sums]
      2. rank totals exchanged through peer memory                       [smcb_xchg_allgather]
      3. every rank derives the contiguous range of GLOBAL output slots whose (sorted) uniform falls
         into its CDF segment, searches its own CDF and writes each selected row straight into the
         SoA buffer of the rank that holds the slot over NVLink          [smcb_resample_fused]
      4. distinct-ancestor counts summed through peer memory; the same exchange orders every rank's
         row writes before anyone reads its new particles               [smcb_xchg_allgather]
    Falls back to the routed NCCL form when symmetric memory is unavailable."""
    rank, world = engine.dist_info()
    rng = kernel.rng
    kind = getattr(resample_rng, "kind", None)
    if isinstance(rng, np.random.Generator) or kind is None:
        raise NotImplementedError("multi-GPU resampling needs the native Philox generator and a built-in "
                                  "resampler (replay parity mode is single-GPU)")
    px = engine.PeerExchange.get()
    if px is None or st.peer_cols_alt is None or st.peer_extra_alt is None or _lib.lib.smcb_get_option(b"no_fused_resample"):
        return _resample_sharded_routed(st, kernel, resample_rng, warn_threshold)
    import ctypes as C
    n = st.n
    s = stream_ptr()
    ph = _Phases()
    stream_id = rng.next_stream()
    ph("rs_scan")
    engine.weights_cdf(st, total_out=st.rs_tot[0:1])
    r = _lib.Resample()
    if kind == _lib.RESAMPLE_STANDARD:
        call("smcb_resample_exponential_sums", st.id_offset, n, st.n_global, rng.seed, stream_id, ptr(st.extra_alt),
             ptr(st.rs_tot[1:2]), ptr(st.ws), s)
        for q in range(world):
            r.t_peer[q] = st.peer_extra_alt[q]
    ph("rs_totals")
    totals = px.gather(st.rs_tot)                         # (world, 2) in rank order, identical on every rank
    ph("rs_fused")
    r.kind, r.world, r.rank, r.n_cols = int(kind), world, rank, int(st.cols.shape[0])
    r.n_local, r.n_global, r.seed, r.stream_id = n, st.n_global, rng.seed, stream_id
    r.cdf, r.totals = st.cdf.data_ptr(), totals.data_ptr()
    r.src, r.src_stride = st.cols.data_ptr(), st.col_stride
    for q in range(world):
        r.dst_peer[q] = st.peer_cols_alt[q]
    r.dst_stride = st.col_stride
    r.idx_out, r.n_unique_out = None, st.counts.data_ptr()
    call("smcb_resample_fused", C.byref(r), ptr(st.ws), s)
    ph("rs_tail")
    st.swap_cols()
    n_unique = px.gather(st.counts[0:1].to(torch.float64), reduce=True)     # also the completion barrier
    st.set_uniform_weights()
    n_global = st.n_global
    if defer:
        # already global: scaled so that the step record's sum over ranks gives it back
        st.deferred.append((n_unique / world, lambda v: engine.warn_collapsed(round(v), n_global, warn_threshold)))
    else:
        engine.warn_collapsed(int(n_unique.item()), n_global, warn_threshold)
    ph(None)
    if engine.PROFILE is not None:
        # rows this rank writes into OTHER ranks' buffers over NVLink: slots whose uniform falls into this rank's
        # CDF segment [O_r, O_r+1) minus the slots it holds itself (profiling only: one extra host read)
        tot = totals[:, 0].cpu().numpy()
        off = np.concatenate([[0.0], np.cumsum(tot)]) / max(float(tot.sum()), 1e-300)
        lo_s, hi_s = off[rank] * st.n_global, off[rank + 1] * st.n_global
        own = max(0.0, min(hi_s, (rank + 1) * n) - max(lo_s, rank * n))
        engine._note_bytes("rs_nvlink_rows", max(0.0, (hi_s - lo_s) - own))
    return None


def _resample_sharded_routed(st, kernel, resample_rng, warn_threshold):
    rank, world = engine.dist_info()
    rng = kernel.rng
    kind = getattr(resample_rng, "kind", None)
    if isinstance(rng, np.random.Generator) or kind is None:
        raise NotImplementedError("multi-GPU resampling needs the native Philox generator and a built-in "
                                  "resampler (replay parity mode is single-GPU)")
    n, d = st.n, st.d
    ncol = int(st.cols.shape[0])            # d + 2 (+1 with a proposal log-pdf column)
    s = stream_ptr()
    dv = st.cols.device
    ph = _Phases()
    # 1-2: local CDF and segment offsets
    ph("rs_scan")
    engine.weights_cdf(st)
    totals = engine._all_gather_rows(st.cdf[n - 1:n]).reshape(-1)
    off = segment_offsets(totals)
    # 3: uniforms of my output slots, routed to the owning rank
    ph("rs_uniforms")
    stream_id = rng.next_stream()
    if kind == _lib.RESAMPLE_STANDARD:
        # multinomial: the N i.i.d. uniforms in sorted order by exponential spacings (no sort; slot
        # order is immaterial, particles are exchangeable) -- running sums of per-slot exponentials,
        # scanned like the weights, divided by the grand total
        u = sorted_uniforms(st, rng.seed, stream_id)
    else:
        call("smcb_resample_uniforms", int(kind), st.n_global, st.id_offset, n, rng.seed, stream_id, ptr(st.u), s)
        u = st.u
    ph("rs_route")
    bounds = torch.cat([off[1:world], torch.full((1,), float("inf"), dtype=torch.float64, device=dv)])
    send_counts = route_counts_sorted(u, bounds)
    counts = torch.empty((world, world), dtype=torch.int64, device=dv)      # counts[q][r]: requests q -> r
    dist.all_gather_into_tensor(counts, send_counts.contiguous())
    cm = counts.cpu().numpy()                                               # host sync: split sizes
    sc, rc = cm[rank, :].tolist(), cm[:, rank].tolist()
    n_req = int(sum(rc))
    ph("rs_xchg_u")
    u_in = torch.empty(max(n_req, 1), dtype=torch.float64, device=dv)
    dist.all_to_all_single(u_in[:n_req], u.contiguous(), output_split_sizes=rc, input_split_sizes=sc)
    ph("rs_search")
    n_unique = torch.zeros(1, dtype=torch.int64, device=dv)
    idx = None
    if n_req > 0:
        idx = torch.empty(n_req, dtype=torch.int64, device=dv)
        call("smcb_searchsorted_offset", ptr(st.cdf), n, ptr(off[rank:rank + 1]), ptr(u_in), n_req, ptr(idx), s)
        call("smcb_count_unique", ptr(idx), n_req, n, ptr(st.counts), ptr(st.ws), s)
        n_unique = st.counts[0:1].clone()
    ph("rs_gather")
    if st.peer_cols_alt is not None:
        # 4-5 fused: the owner writes every requested row straight into the requester's SoA
        # buffer over NVLink peer memory (no staging transposes, no NCCL payload exchange)
        block_off = np.concatenate([[0], np.cumsum(cm[:, rank])])               # my incoming blocks by requester
        slot_off = np.array([cm[q, :rank].sum() for q in range(world)])         # where my block sits at requester q
        if n_req > 0:
            import ctypes as C
            dst = (C.c_void_p * world)(*st.peer_cols_alt)
            # keep both offset tensors alive across the launch: temporaries would be returned to
            # the caching allocator at once and alias each other
            bo_dev = engine.dev(block_off, torch.int64)
            so_dev = engine.dev(slot_off, torch.int64)
            call("smcb_gather_p2p", ptr(st.cols), st.col_stride, ptr(idx), n_req, ncol, world,
                 ptr(bo_dev), ptr(so_dev), dst, st.col_stride, s)
        if os.environ.get("SMCB_P2P_VERIFY"):
            torch.cuda.synchronize(); dist.barrier()
            rows_out = torch.empty((max(n_req, 1), ncol), dtype=torch.float64, device=dv)
            soa = torch.empty((ncol, max(n_req, 1)), dtype=torch.float64, device=dv)
            call("smcb_gather", ptr(st.cols), n, ptr(soa), n_req, ptr(idx), n_req, ncol, s)
            call("smcb_soa_to_aos", ptr(soa), ptr(rows_out), n_req, ncol, n_req, s)
            rows_in = torch.empty((n, ncol), dtype=torch.float64, device=dv)
            dist.all_to_all_single(rows_in, rows_out[:n_req], output_split_sizes=sc, input_split_sizes=rc)
            ref = rows_in.t().contiguous()
            bad = (ref != st.cols_alt).sum().item()
            print("[verify] rank", rank, "n_req", n_req, "sc", sc, "rc", rc, "block_off", block_off.tolist(),
                  "slot_off", slot_off.tolist(), "mismatching entries", bad, "of", ref.numel(), flush=True)
        st.swap_cols()      # the all-reduce of n_unique below orders every rank's peer writes before any use
    else:
        # 4: owner-side gather (rows in AoS for the exchange); 5: rows back to the requesters
        rows_out = torch.empty((max(n_req, 1), ncol), dtype=torch.float64, device=dv)
        if n_req > 0:
            soa = torch.empty((ncol, n_req), dtype=torch.float64, device=dv)
            call("smcb_gather", ptr(st.cols), n, ptr(soa), n_req, ptr(idx), n_req, ncol, s)
            call("smcb_soa_to_aos", ptr(soa), ptr(rows_out), n_req, ncol, n_req, s)
        rows_in = torch.empty((n, ncol), dtype=torch.float64, device=dv)
        dist.all_to_all_single(rows_in, rows_out[:n_req], output_split_sizes=sc, input_split_sizes=rc)
        call("smcb_aos_to_soa", ptr(rows_in), ptr(st.cols_alt), n, ncol, n, s)
        st.swap_cols()
    ph("rs_tail")
    st.set_uniform_weights()
    engine._all_reduce_sum(n_unique)
    if int(n_unique.item()) <= max(1, warn_threshold * st.n_global):
        warnings.warn(f"Resampled to less than {warn_threshold * 100}% of particles; ", UserWarning)
    ph(None)
    return None
