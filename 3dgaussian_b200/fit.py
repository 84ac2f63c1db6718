"""On-device multi-view fit loop (the hot loop of the reference's
python/fit_multiview_stub.py:265-311 without the per-view Python autograd graph).

One FitDriver per process / GPU.  Parameters, gradients and Adam state are flat fp32
buffers laid out [means 3N | scales_raw 3N | opacities_raw N | colours C*N]; activations
(softplus+1e-3, sigmoid; fit_multiview_stub.py:268-275) are applied inside the preprocess
kernels, the per-view loss (mean|pred-tgt| + w_sil*mean|alpha-mask| + w_depth*mean|depth/max-d_gt|,
:292-303) and the Adam step with the regulariser gradients (:307-311) are fused kernels of libb2splat.

Multi-GPU (torch.distributed): parameters are replicated, view i belongs to rank i % world;
every rank accumulates the gradients of its views into its flat buffer.  Default tail
(comm="multimem"): the buffers are symmetric-memory allocations bound to an NVLink multicast
address and ONE fused kernel per parameter slice pulls the sum of the owner's share through the
NVSwitch (multimem.ld_reduce), applies the guarded Adam step with the owner's moments and
multicasts the new parameters (multimem.st) -- b2s_adam_step_multimem, pipelined per Gaussian
chunk behind the chain rule.  Fallback (comm="nccl"): one NCCL all-reduce (sum) per chunk -- the
iteration's loss and its pair-buffer overflow count ride in the tail of the same buffer -- then
every rank runs the identical Adam step.  Either way the replicas stay bit-identical without a
broadcast.

Overflow safety: the pair buffers are sized from the parameters at plan() time; Gaussians grow
during a fit.  A view that overflows renders nothing, so its iteration must not reach the
parameters: the Adam kernel is guarded on the device by the (all-reduced) overflow count
(b2s_adam_step_guarded), and the host polls a device counter of skipped steps every
`overflow_check_every` steps, re-plans the buffers and repeats the skipped iterations.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import capi
from .renderer import _ptr, _stream, count_pairs


def local_views(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin view sharding: orbit neighbours (similar cost) land on different ranks."""
    return [i for i in range(num_views) if i % world == rank]


import contextlib
import os

_NVTX = os.environ.get("B2S_NVTX", "") == "1"


@contextlib.contextmanager
def _nvtx(name: str):
    """NVTX range around a phase of the iteration when B2S_NVTX=1 (for nsys / ncu --nvtx timelines); free otherwise."""
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def gaussian_chunks(n: int, c: int) -> List[Tuple[int, int]]:
    """[(first, count)] ranges of the pipelined multi-GPU tail: at most c ranges (fewer for small n: a range holds at least
    4096 Gaussians), every range but the last a multiple of 64 Gaussians, so that every slice of a chunk starts 256-byte
    aligned in the parameter buffers."""
    c = max(1, min(int(c), (n + 4095) // 4096))
    per = ((n + c - 1) // c + 63) // 64 * 64
    return [(i, min(per, n - i)) for i in range(0, n, per)] or [(0, 0)]


def chunk_major_offsets(chunks, sh: int, tail: int = 64) -> Tuple[List[List[int]], int]:
    """Layout of the chunk-major gradient buffer [tail | chunk 0: means, scales, opacities, colours | chunk 1: ...] in
    floats: per chunk the four segment offsets + the chunk's end, segments padded to 64 floats; returns (offsets, total)."""
    al = lambda x: (x + 63) // 64 * 64
    out, o = [], tail
    for first, cnt in chunks:
        segs = []
        for k in (3, 3, 1, 3 * sh):
            segs.append(o)
            o += al(k * cnt)
        out.append(segs + [o])
    return out, o


class FitDriver:
    def __init__(self, n: int, sh_coeffs: int, width: int, height: int,
                 cameras: Sequence[Tuple[Sequence[float], Sequence[float]]], device: torch.device,
                 lr: float = 0.02, silhouette_weight: float = 0.2, reg_opacity: float = 1e-3,
                 reg_scale: float = 1e-3, cutoff_sigma: Optional[float] = None, background=(0.0, 0.0, 0.0),
                 rank: int = 0, world: int = 1, process_group=None, pair_slack: float = 1.25, lanes: int = 1,
                 fused_loss: bool = True, batched_preprocess: bool = True, prepared_budget_bytes: int = 32 << 30,
                 view_groups: int = 1, use_depth: bool = False, depth_weight: float = 0.05,
                 overflow_check_every: int = 8, grad_chunks: Optional[int] = None, comm: Optional[str] = None):
        """use_depth: the fit carries the reference's depth term (fit_multiview_stub.py:298-303, weight
        `depth_weight`): the forward keeps the depth plane and the cutoff defaults to 7 sigma instead of 5, because
        depth = D/(W+1e-6) and its gradient amplify the truncated tails.
        comm (several GPUs): "multimem" = the tail of the iteration is ONE fused kernel per chunk slice -- gradient
        reduce-scatter through the NVSwitch (multimem.ld_reduce), Adam on the owned share, parameter all-gather
        (multimem.st) -- over symmetric-memory buffers (b2s_adam_step_multimem); "nccl" = NCCL all-reduce per chunk +
        replicated Adam.  Default: multimem when NVLink multicast is available on every rank, else nccl
        (B2S_COMM overrides)."""
        if device.type != "cuda":
            raise RuntimeError("FitDriver needs a CUDA device (no CPU fallback)")
        self.n, self.sh, self.W, self.H = int(n), int(sh_coeffs), int(width), int(height)
        self.dev = device
        self.use_depth, self.w_depth = bool(use_depth), float(depth_weight)
        if cutoff_sigma is None:
            cutoff_sigma = 7.0 if self.use_depth else 5.0
        if self.use_depth and not fused_loss:
            raise ValueError("the depth term is part of the fused loss path (fused_loss=True)")
        self.overflow_check_every = max(1, int(overflow_check_every))
        self.lr, self.w_sil = float(lr), float(silhouette_weight)
        self.reg_op, self.reg_scale = float(reg_opacity), float(reg_scale)
        self.rank, self.world, self.pg = rank, world, process_group
        self.num_views = len(cameras)
        self.views = local_views(self.num_views, rank, world)
        act = capi.ACT_SCALES_SOFTPLUS | capi.ACT_OPACITY_SIGMOID | (capi.ACT_COLORS_SIGMOID if self.sh == 1 else 0)
        self.params_c = {i: capi.make_params(width, height, cameras[i][0], cameras[i][1], background,
                                             mode=capi.MODE_WSUM, style=capi.STYLE_TORCH,
                                             cutoff_sigma=cutoff_sigma, sh_coeffs=self.sh, sort_depth=0,
                                             act_flags=act, keep_depth=1 if self.use_depth else 0) for i in self.views}
        # Tail of a multi-GPU iteration: the chain rule, the gradient all-reduce and Adam run as a 3-stage pipeline over
        # `grad_chunks` ranges of Gaussians (chain rule of chunk k+1 | NCCL all-reduce of chunk k on a comm stream | Adam
        # of chunk k-1), instead of three serial full-size steps.  One chunk on a single GPU (nothing to overlap).
        self.view_groups = max(1, int(view_groups))
        self.grad_chunks = max(1, int(grad_chunks)) if grad_chunks is not None else (4 if world > 1 else 1)
        self._force_chunks = grad_chunks is not None and world == 1      # tests: the chunked tail on one GPU
        self._comm_stream = None
        import os as _os
        self.comm = (comm or _os.environ.get("B2S_COMM") or "multimem").lower()
        if self.comm not in ("multimem", "nccl"):
            raise ValueError("comm must be 'multimem' or 'nccl'")
        self._symm = None                        # (parameter handle, gradient handle) of the symmetric buffers
        self._moments_sharded = False            # multimem: m / v of a share are current on its owner only
        self._layout(self.n)
        self.step_no = 0
        self.skipped_dev = torch.zeros(1, dtype=torch.int32, device=device)   # Adam steps the device guard skipped
        self._skipped_host = torch.zeros(1, dtype=torch.int32).pin_memory()    # its last polled copy
        self._poll_event = None                  # pending non-blocking poll
        self._since_check = 0
        self._overflowed = False
        self.profile = None                      # {name: [events]} when bench.py asks for a per-step timeline
        # View lanes: lane l owns its own per-view buffers (images, image gradients, state, workspace, loss and
        # overflow accumulators) and a CUDA stream; view k of this rank runs on lane k % lanes, so the small
        # latency-bound kernels of one view (scans, finalize, loss, g-buffer) overlap the blend kernels of another.
        self.lanes = max(1, int(lanes))
        self.active_lanes = self.lanes            # <= lanes; 1 serialises the views on the caller's stream
        self.fused_loss = bool(fused_loss)
        self.batched_preprocess = bool(batched_preprocess) and self.fused_loss
        self.prepared_budget = int(prepared_budget_bytes)
        self.prepared = None
        self.rgb_l = [torch.empty((height, width, 3), dtype=torch.float32, device=device) for _ in range(self.lanes)]
        self.alpha_l = [torch.empty((height, width), dtype=torch.float32, device=device) for _ in range(self.lanes)]
        self.g_rgb_l = [torch.empty_like(t) for t in self.rgb_l]
        self.g_alpha_l = [torch.empty_like(t) for t in self.alpha_l]
        # per lane {loss sum, overflow count} of the current iteration; summed into the gradient buffer's tail
        self.lane_acc = torch.zeros((self.lanes, 2), dtype=torch.float32, device=device)
        self.loss_l = [self.lane_acc[l, 0:1] for l in range(self.lanes)]
        self.rgb, self.alpha = self.rgb_l[0], self.alpha_l[0]
        self.pair_slack = pair_slack
        self.max_pairs = 0
        self.state = self.ws = None
        self.state_l = self.ws_l = None
        self._lane_streams = None
        self.targets: dict = {}
        self.masks: dict = {}
        self.depths: dict = {}
        # per-view constant blocks for the multi-view chain rule, uploaded once (cameras are fixed)
        L = capi.lib()
        vb = L.b2s_view_block_bytes()
        nv = len(self.views)
        arr = (capi.Params * max(nv, 1))(*[self.params_c[i] for i in self.views])
        host = (C.c_uint8 * (vb * max(nv, 1)))()
        capi.check(L.b2s_pack_views(arr, nv, host))
        self.views_dev = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(device)
        # compact per-view blend-backward sums: (local views, n, 12) floats = 48 B per Gaussian per view
        self.gacc = torch.empty((max(nv, 1), max(self.n, 1), 12), dtype=torch.float32, device=device)
        self._copy_stream = None
        self._stage = None

    def _layout(self, n: int):
        """Flat fp32 layout [means 3n | scales_raw 3n | opacities_raw n | colours 3K n]; segment starts are
        padded to 64 floats (256 B): the kernels read colours with 16-byte loads.  The gradient buffer carries a
        64-float tail behind the parameters' gradients: [0] = this iteration's loss (sum_i loss_i / V), [1] = its
        pair-buffer overflow count -- one all-reduce moves gradients, loss and the Adam guard together."""
        self.n = int(n)
        c = 3 * self.sh
        al = lambda x: (x + 63) // 64 * 64
        self.o_means = 0
        self.o_scales = al(3 * n)
        self.o_opac = self.o_scales + al(3 * n)
        self.o_colors = self.o_opac + al(n)
        self.count = self.o_colors + al(c * n)
        z = lambda extra=0: torch.zeros(self.count + extra, dtype=torch.float32, device=self.dev)
        self.p = self.g = None                   # release (symmetric) buffers of the previous size first
        self._symm = None
        self._moments_sharded = False
        self.m, self.v = z(), z()
        chunks = self._chunks()
        if len(chunks) == 1:
            # gradients laid out like the parameters, the tail behind them
            self.g = z(64)
            self.tail = self.g[self.count:self.count + 64]
            self._gchunks = None
        else:
            # pipelined tail: the gradient buffer is CHUNK-major -- [tail 64 | chunk 0: means, scales, opac, colours |
            # chunk 1: ... ] -- so that a chunk (chunk 0 together with the tail) crosses NVLink as ONE contiguous
            # all-reduce; parameters and moments keep the segment-major layout (Adam takes a pointer per slice)
            self._gchunks, o = chunk_major_offsets(chunks, self.sh)      # per chunk: four segment offsets + end
            if self.comm == "multimem" and self.world > 1:
                self._alloc_symmetric(o)
            if self._symm is None:
                self.g = torch.zeros(o, dtype=torch.float32, device=self.dev)
            self.tail = self.g[0:64]
        if self.p is None:
            self.p = z()
        self.tail_red = torch.zeros(64, dtype=torch.float32, device=self.dev)    # multimem: the reduced tail (local)
        self.loss_dev = self.tail_red[0:1] if self._symm is not None else self.tail[0:1]

    def _alloc_symmetric(self, g_numel: int):
        """Parameter and gradient buffers as symmetric allocations bound to an NVLink multicast address (torch
        symmetric memory does the allocation, the handle exchange and the cross-rank barriers; the kernels on them are
        ours).  All ranks take the same decision: a failure anywhere falls back to NCCL everywhere."""
        ok, err = 1, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.pg if self.pg is not None else torch.distributed.group.WORLD
            p = symm_mem.empty(self.count, dtype=torch.float32, device=self.dev)
            g = symm_mem.empty(g_numel, dtype=torch.float32, device=self.dev)
            p.zero_()
            g.zero_()
            hp, hg = symm_mem.rendezvous(p, group=grp), symm_mem.rendezvous(g, group=grp)
            if not getattr(hp, "multicast_ptr", 0) or not getattr(hg, "multicast_ptr", 0):
                raise RuntimeError("no NVLink multicast address for the symmetric buffers")
        except Exception as e:  # noqa: BLE001
            ok, err = 0, e
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=self.pg)
        if int(flag.item()) == 1:
            self.p, self.g, self._symm = p, g, (hp, hg)
        else:
            self.comm_fallback = str(err) if err is not None else "another rank has no multicast support"

    def _sync_moments(self):
        """multimem mode keeps the Adam moments of a share on its owner only; before anything that reads or permutes
        the whole m / v (checkpoint, spatial reorder) the owners' values are gathered on every rank."""
        if not self._moments_sharded or self._symm is None:
            return
        rank, world = self._symm[0].rank, self._symm[0].world_size
        for buf in (self.m, self.v):
            own = torch.zeros_like(buf)
            for first, count in self._chunks():
                for sl_src, sl_dst, k in zip(self._chunk_slices(buf, first, count), self._chunk_slices(own, first, count),
                                             (3, 3, 1, 3 * self.sh)):
                    lo, hi = capi.multimem_share(k * count, rank, world)      # the kernel's own partition
                    sl_dst[lo:hi] = sl_src[lo:hi]
            torch.distributed.all_reduce(own, group=self.pg)
            buf.copy_(own)
        self._moments_sharded = False

    # ---- parameter views -------------------------------------------------------------------
    def _seg(self, buf, off, numel): return buf[off:off + numel]
    def means(self): return self._seg(self.p, self.o_means, 3 * self.n).view(self.n, 3)
    def scales_raw(self): return self._seg(self.p, self.o_scales, 3 * self.n).view(self.n, 3)
    def opacities_raw(self): return self._seg(self.p, self.o_opac, self.n)
    def colors_raw(self):
        t = self._seg(self.p, self.o_colors, 3 * self.sh * self.n)
        return t.view(self.n, 3) if self.sh == 1 else t.view(self.n, self.sh, 3)

    def grad_views(self):
        """(means, scales_raw, opacities_raw, colours) gradients as flat tensors (views of the gradient buffer, or
        gathered from its chunks when the tail is pipelined).  With comm="multimem" on several GPUs the buffer holds THIS
        rank's contribution only: the sum over the ranks is formed inside the NVSwitch on its way into the owner's Adam
        update and is never written back."""
        n = self.n
        if self._gchunks is None:
            return (self._seg(self.g, self.o_means, 3 * n), self._seg(self.g, self.o_scales, 3 * n),
                    self._seg(self.g, self.o_opac, n), self._seg(self.g, self.o_colors, 3 * self.sh * n))
        out = []
        for q, k in enumerate((3, 3, 1, 3 * self.sh)):
            out.append(torch.cat([self.g[segs[q]:segs[q] + k * cnt] for (first, cnt), segs in zip(self._chunks(), self._gchunks)]))
        return tuple(out)

    def set_params(self, means, scales_raw, opacities_raw, colors_raw):
        with torch.no_grad():
            self.means().copy_(means)
            self.scales_raw().copy_(scales_raw)
            self.opacities_raw().copy_(opacities_raw)
            self.colors_raw().copy_(colors_raw)
        self.m.zero_(); self.v.zero_()
        self.step_no = 0

    def reorder_spatial(self, bits: int = 10) -> torch.Tensor:
        """Stores the Gaussians -- parameters and Adam moments -- in 3-D Morton order of their means and returns the
        permutation (new index -> old index).  The weighted sum is order independent, so the fit is unchanged; what
        changes is locality: Gaussians that are neighbours in space are neighbours on screen in EVERY view, so the
        slice of Gaussians one binning block owns touches a few hundred tiles instead of all of them (its scattered
        4-byte list writes become runs), and the records / gradient rows a tile gathers share cache lines.  A fit calls
        it after initialisation and after densify/prune (the reference re-creates every tensor there anyway,
        python/fit_multiview_stub.py:318-325); it is plain setup work, outside the iteration."""
        n = self.n
        if n <= 1:
            return torch.arange(n, device=self.dev)
        self._sync_moments()
        with torch.no_grad():
            m = self.means()
            lo, hi = m.min(0).values, m.max(0).values
            q = ((m - lo) / (hi - lo).clamp_min(1e-20) * float((1 << bits) - 1)).to(torch.int64).clamp_(0, (1 << bits) - 1)
            code = torch.zeros(n, dtype=torch.int64, device=self.dev)
            for b in range(bits):
                for a in range(3):
                    code |= ((q[:, a] >> b) & 1) << (3 * b + a)
            perm = torch.argsort(code, stable=True)
            segs = ((self.o_means, 3), (self.o_scales, 3), (self.o_opac, 1), (self.o_colors, 3 * self.sh))
            for buf in (self.p, self.m, self.v):
                for off, k in segs:
                    seg = buf[off:off + k * n].view(n, k)
                    seg.copy_(seg[perm])
        return perm

    def _pp(self, off):  # raw device pointer into a flat buffer
        return C.c_void_p(self.p.data_ptr() + 4 * off)

    def _gp(self, off):
        return C.c_void_p(self.g.data_ptr() + 4 * off)

    # ---- capacity ----------------------------------------------------------------------------
    def plan(self, extra_slack: float = 1.0):
        """Sizes the pair buffers from the current parameters (one count pass per local view;
        synchronises).  Called once up front and again when the device guard reports an overflow."""
        worst = 0
        with torch.cuda.device(self.dev):
            sc = torch.nn.functional.softplus(self.scales_raw()) + 1e-3
            op = torch.sigmoid(self.opacities_raw())
            for i in self.views:
                pc = capi.make_params(self.W, self.H, list(self.params_c[i].view), list(self.params_c[i].proj),
                                      (0, 0, 0), cutoff_sigma=self.params_c[i].cutoff_sigma)
                worst = max(worst, count_pairs(pc, self.means().contiguous(), sc.contiguous(), op.contiguous()))
            self.max_pairs = max(int(worst * self.pair_slack * extra_slack) + 4096, 4096)
            L = capi.lib()
            self.state_bytes = L.b2s_state_bytes(self.n, self.W, self.H, self.max_pairs)
            self.ws_bytes = L.b2s_workspace_bytes(self.n, self.W, self.H, self.max_pairs)
            self.state = self.ws = self.state_l = self.ws_l = self.prepared = None      # release before re-allocating
            self.state_l = [torch.empty(self.state_bytes, dtype=torch.uint8, device=self.dev) for _ in range(self.lanes)]
            self.ws_l = [torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev) for _ in range(self.lanes)]
            self._counters_l = [st[:16].view(torch.int32) for st in self.state_l]   # needed(lo,hi), kept, overflow
            self.state, self.ws = self.state_l[0], self.ws_l[0]
            # batched preprocess: one block per local view (b2s_preprocess_views), if it fits the budget and the
            # image has few enough tiles for the counting-sort path the prepared views feed (bin.cu: the tile scan
            # stages 8 B per tile in 200 KB of shared memory); larger images take the per-view radix path
            self.pv_bytes = int(L.b2s_prepared_view_bytes(self.n))
            n_tiles = ((self.W + capi.TILE - 1) // capi.TILE) * ((self.H + capi.TILE - 1) // capi.TILE)
            if (self.batched_preprocess and self.views and n_tiles * 8 <= 200 * 1024 and
                    self.pv_bytes * len(self.views) <= self.prepared_budget):
                self.prepared = torch.empty(self.pv_bytes * len(self.views), dtype=torch.uint8, device=self.dev)
        return worst

    # ---- targets -------------------------------------------------------------------------------
    def set_targets(self, targets: dict, masks: Optional[dict] = None, depths: Optional[dict] = None):
        """Device-resident targets {view index: (H,W,3) float32} (+ masks {(H,W)}, + normalised depth maps {(H,W)}:
        the d_gt of fit_multiview_stub.py:298-303, used when the driver was built with use_depth=True)."""
        if depths and not self.use_depth:
            raise ValueError("depth maps given to a FitDriver built without use_depth=True (the forward would not keep "
                             "the depth plane); construct it with use_depth=True")
        self.targets, self.masks, self.depths = dict(targets), dict(masks or {}), dict(depths or {})

    def render_view(self, i: int, out_rgb=None, out_alpha=None):
        """Forward only (used to synthesise targets).  Re-plans and renders again if the view overflowed."""
        if self.state is None:
            self.plan()
        out_rgb = self.rgb if out_rgb is None else out_rgb
        out_alpha = self.alpha if out_alpha is None else out_alpha
        with torch.cuda.device(self.dev):
            for attempt in range(3):
                capi.check(capi.lib().b2s_forward(
                    capi.ctx(self.dev.index), C.byref(self.params_c[i]), self._pp(self.o_means), self._pp(self.o_scales),
                    self._pp(self.o_colors), self._pp(self.o_opac), self.n, self.max_pairs, _ptr(out_rgb), _ptr(out_alpha),
                    None, _ptr(self.state), self.state_bytes, _ptr(self.ws), self.ws_bytes, _stream()))
                if not capi.ticket_info(self.dev.index)[2]:
                    break
                self.plan(extra_slack=1.5 * (attempt + 1))
            else:
                raise capi.B2SError("render_view: pair buffers overflowed repeatedly")
        return out_rgb, out_alpha

    # ---- one fit iteration -------------------------------------------------------------------
    def _view_fwd_bwd(self, slot: int, i: int, tgt: torch.Tensor, mask: Optional[torch.Tensor], lane: int = 0,
                      ready: Optional[torch.cuda.Event] = None, depth: Optional[torch.Tensor] = None, convert=None):
        """forward + loss + blend backward of view i on the CURRENT stream with lane `lane`'s buffers.  `ready`: event
        after which tgt / mask / depth hold this view's data (host-fed steps): only the loss needs them, so the stream
        waits for it between the forward and the backward, not in front of the forward.  `convert`: called on this
        stream behind `ready` (8-bit host feeds: bytes -> float32 here, so that the copy stream carries DMA only)."""
        L, ctx, st = capi.lib(), capi.ctx(self.dev.index), _stream()
        pc = C.byref(self.params_c[i])
        rgb, alpha, g_rgb, g_alpha = self.rgb_l[lane], self.alpha_l[lane], self.g_rgb_l[lane], self.g_alpha_l[lane]
        state, ws = self.state_l[lane], self.ws_l[lane]
        if self.fused_loss:
            # accumulators only: the loss and its image gradients are evaluated inside the g-buffer kernel
            pv = None
            if self.prepared is not None:
                pv = C.c_void_p(self.prepared.data_ptr() + slot * self.pv_bytes)
                capi.check(L.b2s_forward_prepared(ctx, pc, pv, self.n, self.max_pairs, None, None, None,
                                                  _ptr(state), self.state_bytes, _ptr(ws), self.ws_bytes, st))
            else:
                capi.check(L.b2s_forward(ctx, pc, self._pp(self.o_means), self._pp(self.o_scales),
                                         self._pp(self.o_colors), self._pp(self.o_opac), self.n, self.max_pairs, None,
                                         None, None, _ptr(state), self.state_bytes, _ptr(ws), self.ws_bytes, st))
            self.lane_acc[lane, 1] += self._counters_l[lane][3]
            if ready is not None:
                torch.cuda.current_stream().wait_event(ready)
            if convert is not None:
                convert()
            # 8-bit targets / masks (host-fed steps): the loss kernel reads the image bytes itself
            entry = L.b2s_fit_backward_blend_u8 if tgt.dtype == torch.uint8 else L.b2s_fit_backward_blend
            if mask is not None and mask.dtype != tgt.dtype:
                raise TypeError("target and mask must both be float32 or both be uint8")
            capi.check(entry(ctx, pc, self.n, self.max_pairs, _ptr(tgt), _ptr(mask),
                                                _ptr(depth) if self.use_depth else None, self.w_sil,
                                                self.w_depth if (self.use_depth and depth is not None) else 0.0,
                                                1.0 / self.num_views, _ptr(self.loss_l[lane]), _ptr(state), pv,
                                                _ptr(ws), self.ws_bytes, _ptr(self.gacc[slot]), st))
            return
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
        if convert is not None:
            convert()
        capi.check(L.b2s_forward(ctx, pc, self._pp(self.o_means), self._pp(self.o_scales), self._pp(self.o_colors),
                                 self._pp(self.o_opac), self.n, self.max_pairs, _ptr(rgb), _ptr(alpha), None,
                                 _ptr(state), self.state_bytes, _ptr(ws), self.ws_bytes, st))
        self.lane_acc[lane, 1] += self._counters_l[lane][3]
        capi.check(L.b2s_fit_loss(ctx, _ptr(rgb), _ptr(alpha), _ptr(tgt), _ptr(mask), self.W, self.H,
                                  self.w_sil, 1.0 / self.num_views, _ptr(g_rgb),
                                  _ptr(g_alpha) if mask is not None else None, _ptr(self.loss_l[lane]), st))
        capi.check(L.b2s_backward_blend(ctx, pc, self.n, self.max_pairs, _ptr(g_rgb),
                                        _ptr(g_alpha) if mask is not None else None, None, _ptr(state),
                                        _ptr(ws), self.ws_bytes, _ptr(self.gacc[slot]), st))

    def _preprocess_views(self, a: int, b: int):
        """Per-Gaussian stage of local views [a, b) in one launch (parameters read once per group)."""
        if self.prepared is not None and b > a:
            vb = capi.lib().b2s_view_block_bytes()
            capi.check(capi.lib().b2s_preprocess_views(
                capi.ctx(self.dev.index), C.c_void_p(self.views_dev.data_ptr() + a * vb), b - a, self.sh,
                self._pp(self.o_means), self._pp(self.o_scales), self._pp(self.o_colors), self._pp(self.o_opac), self.n,
                C.c_void_p(self.prepared.data_ptr() + a * self.pv_bytes), _stream()))

    def _chain_rule(self, a: int, b: int, accumulate: bool, first: int = 0, count: Optional[int] = None):
        """Chain rule over local views [a, b) for the Gaussians [first, first + count): every Gaussian's gradients
        summed over the views in registers."""
        vb = capi.lib().b2s_view_block_bytes()
        if self._gchunks is None:
            gm, gs, go, gc = self._gp(self.o_means), self._gp(self.o_scales), self._gp(self.o_opac), self._gp(self.o_colors)
        else:
            # chunk-major gradient buffer: the kernel indexes by the global Gaussian id, so every pointer is the chunk's
            # segment moved back by the chunk's first id
            ci = [f for f, _ in self._chunks()].index(first)
            segs = self._gchunks[ci]
            gm, gs = self._gp(segs[0] - 3 * first), self._gp(segs[1] - 3 * first)
            go, gc = self._gp(segs[2] - first), self._gp(segs[3] - 3 * self.sh * first)
        capi.check(capi.lib().b2s_backward_params_range(
            capi.ctx(self.dev.index), C.c_void_p(self.views_dev.data_ptr() + a * vb), b - a, self.sh,
            self._pp(self.o_means), self._pp(self.o_scales), self._pp(self.o_colors), self._pp(self.o_opac), self.n,
            first, self.n - first if count is None else count,
            _ptr(self.gacc[a]), gm, gs, gc, go, 1 if accumulate else 0, _stream()))

    def _chunks(self):
        """[(first, count)] Gaussian ranges of the pipelined tail (one range when there is nothing to overlap)."""
        c = self.grad_chunks if ((self.world > 1 or self._force_chunks) and self.view_groups == 1) else 1
        return gaussian_chunks(self.n, c)

    def _chunk_slices(self, buf, first, count):
        """The four slices of a flat buffer that belong to the Gaussians [first, first + count)."""
        k = 3 * self.sh
        return [buf[self.o_means + 3 * first: self.o_means + 3 * (first + count)],
                buf[self.o_scales + 3 * first: self.o_scales + 3 * (first + count)],
                buf[self.o_opac + first: self.o_opac + first + count],
                buf[self.o_colors + k * first: self.o_colors + k * (first + count)]]

    def _streams(self):
        if self._lane_streams is None:
            self._lane_streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.lanes)]
            self._tail_stream = torch.cuda.Stream(device=self.dev)
            self._step_begin = torch.cuda.Event()
            self._tail_done = torch.cuda.Event()
        return self._lane_streams

    def _iterate(self, inputs):
        """One pass over this rank's views: forward + loss + blend backward per view, chain rule, on `lanes`
        concurrent streams.  inputs(k, stream) -> dict(tgt=, mask=, depth=, done=callback, ready=event) for local view k.

        With several lanes the views are cut into `view_groups` groups and pipelined: the caller's stream runs the
        batched preprocess of every group (group g+1's while the lanes blend group g), the lanes wait for their
        group's records, and a tail stream folds each finished group into the gradients (accumulating after the
        first) while later groups are still blending -- only the first group's preprocess and the last group's
        chain rule are exposed."""
        main = torch.cuda.current_stream()
        nv = len(self.views)
        nl = max(1, min(self.active_lanes, self.lanes))
        self._mark("step_begin")
        self.lane_acc.zero_()
        self._chain_job = None
        if nv == 0:
            self.g.zero_()
            return

        def one_view(k, lane, stream):
            q = inputs(k, stream)
            if _NVTX:
                torch.cuda.nvtx.mark(f"b2s.view {self.views[k]} lane {lane}")
            self._view_fwd_bwd(k, self.views[k], q["tgt"], q.get("mask"), lane, q.get("ready"), q.get("depth"), q.get("convert"))
            if q.get("done") is not None:
                q["done"](stream)

        chunks = self._chunks()
        if nl == 1:
            self._preprocess_views(0, nv)
            for k in range(nv):
                one_view(k, 0, main)
            torch.sum(self.lane_acc[:nl], dim=0, out=self.tail[0:2])
            if len(chunks) > 1:
                self._chain_job = (main, 0, nv)          # _finish_step interleaves chain rule, all-reduce and Adam per chunk
            else:
                self._chain_rule(0, nv, False)
            return
        else:
            streams = self._streams()
            tail = self._tail_stream
            self._step_begin.record(main)
            for st in streams + [tail]:
                st.wait_event(self._step_begin)       # after everything queued on the caller's stream (the last Adam)
            G = max(1, min(self.view_groups, nv // nl))
            bounds = [(g * nv) // G for g in range(G + 1)]
            # The batched preprocess is cut into two launches, both queued up front on the caller's stream: the first
            # covers one view per lane, so the lanes start blending after a sliver of the head (2.1 ms of a 42 ms
            # iteration on one GPU, 0.3 of 6.5 ms on eight) instead of all of it; the second (every other view) runs
            # beside those blends.  (Eight equal launches measured slower: each re-reads the 220 MB of parameters and
            # takes SMs from the blends.)  With view_groups > 1 a launch is a view group.
            PG = G if G > 1 else (2 if nv > nl else 1)
            pbounds = bounds if G > 1 else ([0, nl, nv] if nv > nl else [0, nv])
            pre_ev = []
            for j in range(PG):
                self._preprocess_views(pbounds[j], pbounds[j + 1])
                ev = torch.cuda.Event()
                ev.record(main)
                pre_ev.append(ev)
            self._mark("preprocess_done")
            pgroup = lambda k: next(j for j in range(PG) if pbounds[j] <= k < pbounds[j + 1])
            waited = [-1] * nl                         # last preprocess launch each lane has waited for
            for g in range(G):
                a, b = bounds[g], bounds[g + 1]
                used = sorted({k % nl for k in range(a, b)})
                for k in range(a, b):
                    lane = k % nl
                    j = pgroup(k)
                    if waited[lane] < j:
                        streams[lane].wait_event(pre_ev[j])
                        waited[lane] = j
                    with torch.cuda.stream(streams[lane]):
                        one_view(k, lane, streams[lane])
                for l in used:
                    ev = torch.cuda.Event()
                    ev.record(streams[l])
                    tail.wait_event(ev)
                with torch.cuda.stream(tail):
                    if g == G - 1:
                        self._mark("views_done", tail)
                        # loss and overflow count of this rank's views into the gradient buffer's tail (fixed order:
                        # deterministic); the tail stream has waited for every lane
                        torch.sum(self.lane_acc[:nl], dim=0, out=self.tail[0:2])
                    if G == 1 and len(chunks) > 1:
                        self._chain_job = (tail, a, b)      # _finish_step interleaves chain rule, all-reduce and Adam per chunk
                    else:
                        self._chain_rule(a, b, g > 0)
            if self._chain_job is None:
                self._mark("chain_done", tail)
                self._tail_done.record(tail)
                main.wait_event(self._tail_done)       # the tail has waited for every lane
            # (chunked tail: the caller's stream waits for each chunk's all-reduce in _finish_step instead, and that
            # all-reduce waits for the chunk's chain rule -- nothing is queued behind the whole chain rule)

    def _adam(self, p, g, m, v, reg: Optional[str], frac: float, count_skip: bool):
        """Guarded Adam over one contiguous slice; reg = 'scales' / 'opac' applies that regulariser to the whole slice
        (its weight scaled by the slice's share `frac` of the segment, so that the per-element gradient stays
        reg / segment size)."""
        nel = p.numel()
        if nel == 0:
            return
        sb = se = ob = oe = 0
        rs = ro = 0.0
        if reg == "scales":
            se, rs = nel, self.reg_scale * frac
        elif reg == "opac":
            oe, ro = nel, self.reg_op * frac
        capi.check(capi.lib().b2s_adam_step_guarded(
            capi.ctx(self.dev.index), _ptr(p), _ptr(g), _ptr(m), _ptr(v), nel, self.step_no, self.lr, 0.9, 0.999, 1e-8,
            sb, se, rs, ob, oe, ro, C.c_void_p(self.tail.data_ptr() + 4), _ptr(self.skipped_dev) if count_skip else None,
            _stream()))

    def _finish_step(self):
        self.step_no += 1
        if len(self._chunks()) > 1:      # the same decision on every rank (an idle rank simply has no chain rule to run)
            # 3-stage pipeline over the Gaussian chunks, queued chunk by chunk so that the GPU sees chunk k's all-reduce
            # BEFORE chunk k+1's chain rule (kernels that arrive first get the SMs first): chain rule on the tail stream |
            # ONE contiguous all-reduce per chunk on the comm stream (the gradient buffer is chunk-major; loss and
            # overflow count -- the Adam guard -- ride with chunk 0) | guarded Adam per slice on the caller's stream.
            if self._comm_stream is None:
                # HIGH priority: when the exchange + Adam of chunk k become ready, the chain-rule grids of chunks k+1.. are
                # already resident or pending, and at equal priority the block scheduler drains those first -- the whole
                # exchange then queued up BEHIND the chain rule instead of beside it (2 GPUs: chunk 0 reduced 0.9 ms after
                # the last view, its Adam done at 1.67 ms, when the last chain rule ended at 1.60 ms)
                self._comm_stream = torch.cuda.Stream(device=self.dev, priority=-1)
            main, comm = torch.cuda.current_stream(), self._comm_stream
            job = getattr(self, "_chain_job", None)
            chunks = self._chunks()
            if self._symm is not None:
                self._finish_step_multimem(main, comm, job, chunks)
                return
            for c, (first, count) in enumerate(chunks):
                segs = self._gchunks[c]
                gs = [self.g[segs[q]:segs[q] + k * count] for q, k in enumerate((3, 3, 1, 3 * self.sh))]
                if job is not None:
                    with torch.cuda.stream(job[0]):
                        self._chain_rule(job[1], job[2], False, first, count)
                        chain_ev = torch.cuda.Event()
                        chain_ev.record(job[0])
                        self._mark(f"chunk{c}_chain", job[0])
                with torch.cuda.stream(comm):
                    if job is not None:
                        comm.wait_event(chain_ev)
                    elif c == 0:
                        comm.wait_stream(main)          # idle rank (no local views): its zeroed buffer is on the caller's stream
                    if self.world > 1:
                        torch.distributed.all_reduce(self.g[0 if c == 0 else segs[0]:segs[4]], group=self.pg)
                    self._mark(f"chunk{c}_ar", comm)
                    # Adam of the chunk on the same high-priority stream, right behind its all-reduce
                    ps, ms, vs = (self._chunk_slices(b, first, count) for b in (self.p, self.m, self.v))
                    frac = count / max(self.n, 1)
                    for q, reg in enumerate((None, "scales", "opac", None)):
                        self._adam(ps[q], gs[q], ms[q], vs[q], reg, frac, count_skip=(c == 0 and q == 0))
                    ev = torch.cuda.Event()
                    ev.record(comm)
                    self._mark(f"chunk{c}_adam", comm)
                main.wait_event(ev)
            self._chain_job = None
            self._mark("adam_done")
            return
        if self.world > 1:
            torch.distributed.all_reduce(self.g, group=self.pg)     # gradients + loss + overflow count: ONE collective
        self._mark("allreduce_done")
        capi.check(capi.lib().b2s_adam_step_guarded(
            capi.ctx(self.dev.index), _ptr(self.p), _ptr(self.g), _ptr(self.m), _ptr(self.v), self.count, self.step_no,
            self.lr, 0.9, 0.999, 1e-8, self.o_scales, self.o_scales + 3 * self.n, self.reg_scale, self.o_opac,
            self.o_opac + self.n, self.reg_op, C.c_void_p(self.tail.data_ptr() + 4), _ptr(self.skipped_dev), _stream()))
        self._mark("adam_done")

    def _finish_step_multimem(self, main, comm, job, chunks):
        """The tail over NVLink multicast: per chunk  chain rule (tail stream) | barrier, fused reduce-scatter + Adam +
        all-gather of the chunk's four slices, barrier (comm stream).  No NCCL call, no separate Adam pass; the caller's
        stream waits for the last chunk's second barrier, i.e. for every rank's share of the new parameters."""
        hp, hg = self._symm
        L, ctx = capi.lib(), capi.ctx(self.dev.index)
        mc_p, mc_g = int(hp.multicast_ptr), int(hg.multicast_ptr)
        rank, world = int(hp.rank), int(hp.world_size)
        ks = (3, 3, 1, 3 * self.sh)
        for c, (first, count) in enumerate(chunks):
            segs = self._gchunks[c]
            p_offs = (self.o_means + 3 * first, self.o_scales + 3 * first, self.o_opac + first, self.o_colors + 3 * self.sh * first)
            if job is not None:
                with torch.cuda.stream(job[0]):
                    self._chain_rule(job[1], job[2], False, first, count)
                    chain_ev = torch.cuda.Event()
                    chain_ev.record(job[0])
                    self._mark(f"chunk{c}_chain", job[0])
            with torch.cuda.stream(comm):
                if job is not None:
                    comm.wait_event(chain_ev)
                elif c == 0:
                    comm.wait_stream(main)              # idle rank (no local views): its zeroed buffer is on the caller's stream
                hg.barrier(channel=0)                   # every rank's gradients of this chunk are written
                st = _stream()
                if c == 0:                              # loss + overflow count (the Adam guard), summed by the switch
                    capi.check(L.b2s_reduce_tail_multimem(ctx, C.c_void_p(mc_g), _ptr(self.tail_red), 2, st))
                frac = count / max(self.n, 1)
                for q, reg in enumerate((None, "scales", "opac", None)):
                    nel = ks[q] * count
                    if nel == 0:
                        continue
                    se = oe = 0
                    rs = ro = 0.0
                    if reg == "scales":
                        se, rs = nel, self.reg_scale * frac
                    elif reg == "opac":
                        oe, ro = nel, self.reg_op * frac
                    po = 4 * p_offs[q]
                    capi.check(L.b2s_adam_step_multimem(
                        ctx, C.c_void_p(mc_p + po), C.c_void_p(mc_g + 4 * segs[q]), C.c_void_p(self.p.data_ptr() + po),
                        C.c_void_p(self.m.data_ptr() + po), C.c_void_p(self.v.data_ptr() + po), nel, rank, world, self.step_no,
                        self.lr, 0.9, 0.999, 1e-8, 0, se, rs, 0, oe, ro, C.c_void_p(self.tail_red.data_ptr() + 4),
                        _ptr(self.skipped_dev) if (c == 0 and q == 0) else None, st))
                hg.barrier(channel=0)                   # every rank's share of the new parameters has landed everywhere
                ev = torch.cuda.Event()
                ev.record(comm)
                self._mark(f"chunk{c}_ar", comm)
                self._mark(f"chunk{c}_adam", comm)
            main.wait_event(ev)
        self._moments_sharded = True
        self._chain_job = None
        self._mark("adam_done")

    def _device_inputs(self, k, stream):
        i = self.views[k]
        return {"tgt": self.targets[i], "mask": self.masks.get(i), "depth": self.depths.get(i)}

    def _poll_overflow(self):
        """Non-blocking poll of the guard's skipped-step counter: every `overflow_check_every` steps the counter is
        copied to pinned host memory behind the step (no wait); a later step looks at the copy once its event has
        completed.  The host never drains the stream for it -- steps queued behind an overflow are skipped by the same
        guard (unchanged parameters overflow again), so reacting a few steps late loses nothing."""
        if self._poll_event is not None and self._poll_event.query():
            self._poll_event = None
            if int(self._skipped_host[0]) != 0:
                self._resolve_overflow()
                return
        self._since_check += 1
        if self._poll_event is None and self._since_check >= self.overflow_check_every:
            self._since_check = 0
            self._skipped_host.copy_(self.skipped_dev, non_blocking=True)
            self._poll_event = torch.cuda.Event()
            self._poll_event.record()

    def _mark(self, name, stream=None):
        if self.profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else torch.cuda.current_stream())
            self.profile.setdefault(name, []).append(ev)

    def _resolve_overflow(self) -> int:
        """Reads the device counter of Adam steps the overflow guard skipped (synchronises).  If there were any, every
        step since the first one was skipped too (unchanged parameters overflow again), so the host step count is
        rolled back, the pair buffers are re-planned from the current parameters with more slack, and the skipped
        iterations are repeated.  Returns the number of iterations that had to be repeated."""
        self._since_check = 0
        self._poll_event = None
        k = int(self.skipped_dev.item())
        if k == 0:
            return 0
        self._overflowed = True
        self.skipped_dev.zero_()
        self.step_no -= k
        self.plan(extra_slack=1.5)
        for _ in range(k):
            self._iterate(self._device_inputs)
            self._finish_step()
        if int(self.skipped_dev.item()) != 0:
            raise capi.B2SError("pair buffers overflowed again right after re-planning")
        return k

    def step(self):
        """fwd + bwd over this rank's views (device-resident targets) + all-reduce + Adam.
        Returns the device scalar holding sum_i loss_i / V (without the regulariser).  No host synchronisation,
        except every `overflow_check_every` steps to poll the overflow guard (see the module docstring)."""
        if self.state is None:
            self.plan()
        with torch.cuda.device(self.dev):
            with _nvtx("b2s.views"):            # batched preprocess + forward / loss / blend backward of the local views
                self._iterate(self._device_inputs)
            with _nvtx("b2s.tail"):             # chain rule | gradient exchange | Adam
                self._finish_step()
            self._poll_overflow()
        return self.loss_dev

    def step_from_host(self, host_targets: dict, host_masks: Optional[dict] = None,
                       host_depths: Optional[dict] = None, defer_loss: bool = False):
        """Same iteration fed from PINNED HOST buffers: every view's target (and mask, and depth map) is copied
        host->device inside the step (two staging slots per lane on a side stream, so the copy of a later
        view overlaps the kernels of the current ones) and the loss is read back to the host at the end.
        Targets / masks / depth maps may be float32 in [0,1] or uint8 (decoded image bytes, converted on the device).

        defer_loss=True: the step's (loss, overflow count) still cross to pinned host memory every step, but with a
        non-blocking copy; the call returns the PREVIOUS step's loss (None on the first call) and the host waits for
        that older copy only, so the device always has the next iteration queued when one ends instead of idling
        through the host's wake-up and the first launches (what a training loop that logs its loss asynchronously does).
        `flush_loss()` returns the last step's loss.  An overflow is then seen one step late: both steps were skipped by
        the device guard, the buffers are re-planned and both are repeated."""
        if defer_loss:
            return self._step_from_host_deferred(host_targets, host_masks, host_depths)
        self.flush_loss()
        for attempt in range(3):
            self._queue_host_step(host_targets, host_masks, host_depths, wait_free=attempt > 0)
            with torch.cuda.device(self.dev):
                # the step's one device->host read: loss + overflow guard
                loss, overflow = (self.tail_red if self._symm is not None else self.tail)[0:2].tolist()
            if overflow == 0.0:
                self._since_check = 0
                return float(loss)
            # the guard skipped this Adam step: re-plan with more room and run the iteration again
            self._overflowed = True
            self.skipped_dev.zero_()
            self.step_no -= 1
            self.plan(extra_slack=1.5 * (attempt + 1))
        raise capi.B2SError("pair buffers overflowed repeatedly")

    def _step_from_host_deferred(self, host_targets, host_masks, host_depths):
        if getattr(self, "_loss_ring", None) is None:
            self._loss_ring = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._ring_i = 0
            self._pending_loss = None
        self._queue_host_step(host_targets, host_masks, host_depths, wait_free=True)
        slot = self._loss_ring[self._ring_i & 1]
        self._ring_i += 1
        with torch.cuda.device(self.dev):
            slot.copy_((self.tail_red if self._symm is not None else self.tail)[0:2], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        prev, self._pending_loss = self._pending_loss, (slot, ev, (host_targets, host_masks, host_depths))
        if prev is None:
            return None
        prev[1].synchronize()
        loss, overflow = prev[0].tolist()
        if overflow == 0.0:
            return float(loss)
        # the previous step overflowed its pair buffers: the guard skipped its Adam step, and this one's too (unchanged
        # parameters overflow again).  Drain, re-plan with more room, repeat both with a blocking read.
        ev.synchronize()
        self._pending_loss = None
        self._overflowed = True
        self.skipped_dev.zero_()
        self.step_no -= 2
        self.plan(extra_slack=1.5)
        loss = self.step_from_host(*prev[2])
        self._last_flushed = self.step_from_host(host_targets, host_masks, host_depths)
        return loss

    def flush_loss(self):
        """Loss of the last deferred step (waits for it); None when nothing is pending."""
        pend = getattr(self, "_pending_loss", None)
        if pend is None:
            return getattr(self, "_last_flushed", None)
        self._pending_loss = None
        pend[1].synchronize()
        loss, overflow = pend[0].tolist()
        if overflow != 0.0:
            self._overflowed = True
            self.skipped_dev.zero_()
            self.step_no -= 1
            self.plan(extra_slack=1.5)
            loss = self.step_from_host(*pend[2])
        self._last_flushed = float(loss)
        return self._last_flushed

    def _queue_host_step(self, host_targets, host_masks, host_depths, wait_free: bool):
        """Queues one host-fed iteration (copies, views, tail) without reading anything back.  wait_free: the staging
        slots may still be read by an earlier iteration (a repeated attempt, or a deferred-loss step that did not drain
        the stream), so every copy waits for its slot's last consumer."""
        if self.state is None:
            self.plan()
        if host_depths is not None and not self.use_depth:
            raise ValueError("depth maps given to a FitDriver built without use_depth=True")
        with torch.cuda.device(self.dev):
            nslots = 2 * self.lanes
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.dev)
                self._stage = [(torch.empty_like(self.rgb), torch.empty_like(self.alpha), torch.empty_like(self.alpha))
                               for _ in range(nslots)]
                self._ev_ready = [torch.cuda.Event() for _ in range(nslots)]
                self._ev_free = [torch.cuda.Event() for _ in range(nslots)]
                self._stage_u8 = {}
            use_mask = host_masks is not None
            use_depth = host_depths is not None
            nv = len(self.views)
            pending = {}                                  # slot -> [(raw bytes, float32 destination)] awaiting conversion
            staged = {}                                   # slot -> (target, mask, depth) tensors the view's kernels read
            # With the fused loss, 8-bit targets and masks are consumed as bytes (b2s_fit_backward_blend_u8): no conversion
            # kernel, no float32 copy.  Both must then be bytes; depth maps (and everything on the unfused path) are converted.
            direct_u8 = (self.fused_loss and all(host_targets[i].dtype == torch.uint8 for i in self.views) and
                         (not use_mask or all(host_masks[i].dtype == torch.uint8 for i in self.views)))

            def upload(dst, src, slot, which):
                """host -> stage; 8-bit images cross PCIe as bytes and are converted on the device
                (np.asarray(img, float32) / 255, fit_multiview_stub.py:16-23): targets and masks by the fused loss kernel
                itself (it reads the bytes), anything else by the CONSUMER on the view's own lane stream (`convert`
                below).  A conversion kernel on the copy stream has to find free SM slots between the persistent blend
                kernels of four lanes before the next DMA may start, which made the 8-bit feed slower than the float32
                one that moves four times the bytes (21.9 vs 22.7 iters/s at C4)."""
                if src.dtype == torch.uint8:
                    key = (slot, which)
                    if key not in self._stage_u8:
                        self._stage_u8[key] = torch.empty(src.shape, dtype=torch.uint8, device=self.dev)
                    raw = self._stage_u8[key]
                    raw.copy_(src, non_blocking=True)
                    if direct_u8 and which < 2:
                        return raw                        # target / mask bytes go to the loss kernel as they are
                    pending.setdefault(slot, []).append((raw, dst))
                else:
                    dst.copy_(src, non_blocking=True)
                return dst

            def make_convert(jobs):
                def convert():
                    for raw, dst in jobs:
                        capi.check(capi.lib().b2s_u8_to_f32(capi.ctx(self.dev.index), _ptr(raw), _ptr(dst), raw.numel(),
                                                            _stream()))
                return convert if jobs else None

            issued = [0]

            def issue_upto(k_hi):
                # copies are queued in view order on the copy stream, at most nslots ahead of the consumers
                while issued[0] < min(k_hi, nv):
                    k = issued[0]
                    slot = k % nslots
                    i = self.views[k]
                    with torch.cuda.stream(self._copy_stream):
                        if k >= nslots or wait_free:
                            self._copy_stream.wait_event(self._ev_free[slot])
                        pending[slot] = []
                        t_dev = upload(self._stage[slot][0], host_targets[i], slot, 0)
                        m_dev = upload(self._stage[slot][1], host_masks[i], slot, 1) if use_mask else None
                        d_dev = upload(self._stage[slot][2], host_depths[i], slot, 2) if use_depth else None
                        staged[slot] = (t_dev, m_dev, d_dev)
                        self._ev_ready[slot].record(self._copy_stream)
                    issued[0] += 1

            def inputs(k, st):
                issue_upto(k + self.lanes + 1)       # view k's slot was freed (recorded) before this point
                slot = k % nslots
                t_dev, m_dev, d_dev = staged[slot]
                return {"tgt": t_dev, "mask": m_dev, "depth": d_dev,
                        "done": lambda s, slot=slot: self._ev_free[slot].record(s), "ready": self._ev_ready[slot],
                        "convert": make_convert(list(pending.get(slot, [])))}

            with _nvtx("b2s.views(host-fed)"):
                self._iterate(inputs)
            with _nvtx("b2s.tail"):
                self._finish_step()

    # ---- densify / prune ----------------------------------------------------------------------
    def densify_prune(self, iteration: int, max_gaussians: int, densify_ratio: float = 0.15,
                      prune_opacity: float = 0.05, seed: int = 0, reorder: bool = False) -> int:
        """Device-side _densify_and_prune (reference python/fit_multiview_stub.py:140-197, called every
        densify_prune_interval iterations, :318-325): compacts the survivors, appends the clones, rebuilds
        the flat buffers for the new count and resets the Adam state like the reference's fresh optimizer.
        Philox(seed, iteration, source index) jitter => identical on every rank, no broadcast.
        `reorder=True` re-sorts the new set into 3-D Morton order (the reference order -- survivors, then clones --
        is what `reorder=False` keeps and what the parity tests compare)."""
        L = capi.lib()
        n, cf = self.n, 3 * self.sh
        cap = max(int(max_gaussians), n, 1)
        with torch.cuda.device(self.dev):
            om = torch.empty((cap, 3), dtype=torch.float32, device=self.dev)
            os_ = torch.empty((cap, 3), dtype=torch.float32, device=self.dev)
            oo = torch.empty((cap,), dtype=torch.float32, device=self.dev)
            oc = torch.empty((cap, cf), dtype=torch.float32, device=self.dev)
            wsb = L.b2s_densify_workspace_bytes(n)
            ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
            n_new = C.c_int(0)
            capi.check(L.b2s_densify_prune(capi.ctx(self.dev.index), self._pp(self.o_means), self._pp(self.o_scales),
                                           self._pp(self.o_opac), self._pp(self.o_colors), n, cf, int(max_gaussians),
                                           float(densify_ratio), float(prune_opacity), int(seed), int(iteration),
                                           _ptr(om), _ptr(os_), _ptr(oo), _ptr(oc), C.byref(n_new), _ptr(ws), wsb,
                                           _stream()))
            k = int(n_new.value)
            self._layout(k)                                    # new p/g and zeroed m/v: the Adam reset
            self.step_no = 0
            colors = oc[:k].view(k, 3) if self.sh == 1 else oc[:k].view(k, self.sh, 3)
            with torch.no_grad():
                self.means().copy_(om[:k]); self.scales_raw().copy_(os_[:k])
                self.opacities_raw().copy_(oo[:k]); self.colors_raw().copy_(colors)
            self.gacc = torch.empty((max(len(self.views), 1), max(k, 1), 12), dtype=torch.float32, device=self.dev)
            self.state = self.ws = self.state_l = self.ws_l = self.prepared = None
            if reorder:          # the clones were appended at the end: restore the spatial order (see reorder_spatial)
                self.reorder_spatial()
            self.plan()
        return k

    # ---- checkpoint / resume ---------------------------------------------------------------------
    def save_checkpoint(self, path) -> None:
        """Raw parameters + Adam moments + step count as one .npz (the reference only ever writes the ACTIVATED
        model, python/fit_multiview_stub.py:338-354, and cannot resume; `io.save_gaussians_npz` is that file)."""
        import numpy as np
        n = self.n
        self._sync_moments()
        seg = lambda buf, off, k: buf[off:off + k * n].view(n, k).cpu().numpy()
        arrs = {"format": np.array("b2splat-fit-checkpoint-1"), "n": np.int64(n), "sh_coeffs": np.int64(self.sh),
                "step": np.int64(self.step_no), "lr": np.float64(self.lr)}
        for name, buf in (("p", self.p), ("m", self.m), ("v", self.v)):
            arrs[name + "_means"] = seg(buf, self.o_means, 3)
            arrs[name + "_scales_raw"] = seg(buf, self.o_scales, 3)
            arrs[name + "_opacities_raw"] = seg(buf, self.o_opac, 1)[:, 0]
            arrs[name + "_colors_raw"] = seg(buf, self.o_colors, 3 * self.sh)
        with open(path, "wb") as f:
            np.savez(f, **arrs)

    def load_checkpoint(self, path) -> None:
        """Restores what save_checkpoint wrote; the Gaussian count may differ from the constructor's (a checkpoint
        taken after densify/prune): the flat buffers and the pair buffers are rebuilt for it."""
        import numpy as np
        z = np.load(path, allow_pickle=False)
        if str(z["format"]) != "b2splat-fit-checkpoint-1":
            raise ValueError(f"{path}: not a b2splat fit checkpoint")
        n, sh = int(z["n"]), int(z["sh_coeffs"])
        if sh != self.sh:
            raise ValueError(f"{path}: checkpoint has {sh} colour coefficients per Gaussian, the driver {self.sh}")
        with torch.cuda.device(self.dev):
            if n != self.n:
                self._layout(n)
                self.gacc = torch.empty((max(len(self.views), 1), max(n, 1), 12), dtype=torch.float32, device=self.dev)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.dev)
            with torch.no_grad():
                for name, buf in (("p", self.p), ("m", self.m), ("v", self.v)):
                    for key, off, k in (("means", self.o_means, 3), ("scales_raw", self.o_scales, 3),
                                        ("opacities_raw", self.o_opac, 1), ("colors_raw", self.o_colors, 3 * self.sh)):
                        buf[off:off + k * n].copy_(t(z[f"{name}_{key}"]).reshape(-1))
            self.step_no = int(z["step"])
            self.state = self.ws = self.state_l = self.ws_l = self.prepared = None
            self.plan()

    def check_overflow(self) -> bool:
        """True if any iteration since the last call overflowed its pair buffers.  Such iterations never reached the
        parameters (device guard); by the time this returns they have been repeated with re-planned buffers.
        Synchronises."""
        with torch.cuda.device(self.dev):
            self._resolve_overflow()
        v, self._overflowed = self._overflowed, False
        return v
