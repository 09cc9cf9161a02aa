"""Image-sharded data parallelism for the loss path (the reference itself is single-GPU).

Every quantity of FocalLoss is per image (per-image GT, per-image npos normaliser), so the batch is split into
contiguous image shards, one process per GPU, and the ONLY exchange is an all-gather of the per-image terms
(bg_j, fg_j, reg_j, enhance_j: <= 4*N floats) so that every rank can form exactly the reductions the caller
applies (IL_Loss: .mean() over the batch, or the clip_loss mask on per-image fg, losses.py:575-588).
Gradients of the head outputs never leave the rank that owns the images.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


def shard_sizes(n_global, world):
    """Contiguous shards, remainder spread over the first ranks (same rule on every rank)."""
    base, rem = divmod(int(n_global), int(world))
    return [base + (1 if r < rem else 0) for r in range(world)]


def shard_slice(n_global, world, rank):
    sizes = shard_sizes(n_global, world)
    start = sum(sizes[:rank])
    return slice(start, start + sizes[rank])


class PeerGather:
    """Per-process state of the FUSED all-gather (include/cldet.h, cldet_focal_loss_sharded): a gather buffer
    float[2][world][4][N] plus arrival counters uint32[2][world], allocated by libcldet (cudaMalloc), exported with CUDA IPC
    and opened by every other rank of the group with ITS device current, so that kernels there can store into it over NVLink.
    The loss kernel's last block per image writes the image's four terms straight into every rank's buffer; `wait()` enqueues
    the one-block kernel that blocks the stream until all ranks have delivered and copies the terms into a private tensor.

    Protocol: the arrival counters only ever grow; the u-th use of a parity waits for u*N arrivals per source rank, so a late
    arrival is never mistaken for the next step's.  The wait is bounded (`timeout_ms`, default $CLDET_PEER_TIMEOUT_MS or
    20 000): on a timeout the kernel fills the missing rank's rows with NaN and raises a status word that lives in mapped
    pinned HOST memory; the next `wait()`/`check()` on this rank reads it without a device synchronisation and raises
    CldetError -- a straggler can make a step fail loudly, never silently wrong."""

    def __init__(self, n_local, device, group=None, timeout_ms=None):
        import ctypes
        import os

        from . import _lib
        lib = _lib.load()
        self._lib = _lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = int(n_local)
        self.device = device
        self.timeout_ms = int(timeout_ms if timeout_ms is not None else os.environ.get('CLDET_PEER_TIMEOUT_MS', '20000'))
        n_terms = 2 * self.world * 4 * self.n
        self._flag_off = n_terms + ((-n_terms) % 64)
        numel = self._flag_off + 2 * self.world + 64
        self._own = 0
        self._opened = []
        self.uses = [0, 0]          # how many times each parity has been used (the wait target is uses * N arrivals)
        self.parity = 0
        # ---- local, fallible part: no collective in here ----
        err = None
        hbuf = ctypes.create_string_buffer(64)
        try:
            with torch.cuda.device(device):
                ptr = ctypes.c_void_p()
                _lib.check(lib.cldet_peer_alloc(4 * numel, ctypes.byref(ptr), hbuf))
                self._own = ptr.value
        except Exception as e:  # noqa: BLE001
            err = e
        # ---- collectives: entered by EVERY rank exactly once each, whatever happened locally ----
        handles = [None] * self.world
        dist.all_gather_object(handles, hbuf.raw if err is None else None, group=group)
        base_ptrs = []
        if err is None and all(h is not None for h in handles):
            try:
                with torch.cuda.device(device):
                    for r in range(self.world):
                        if r == self.rank:
                            base_ptrs.append(self._own)
                            continue
                        q = ctypes.c_void_p()
                        _lib.check(lib.cldet_peer_open(handles[r], ctypes.byref(q)))
                        self._opened.append(q.value)
                        base_ptrs.append(q.value)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError('a peer rank could not allocate its exchange buffer')
        ok = torch.tensor([1 if err is None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                       # every rank has zeroed, exported and mapped its buffers (or given up)
        if int(ok.item()) == 0:
            self.close()
            raise _lib.CldetError('peer exchange unavailable on at least one rank%s' % ('' if err is None else ': %r' % (err,)))
        with torch.cuda.device(device):
            self.term_ptrs = torch.tensor(base_ptrs, dtype=torch.int64, device=device)
            self.flag_ptrs = torch.tensor([p + 4 * self._flag_off for p in base_ptrs], dtype=torch.int64, device=device)
        # status word in pinned host memory (UVA: the same address is valid on the device): readable without a sync
        self.status = torch.zeros(1, dtype=torch.int32).pin_memory()

    def exchange(self):
        return self._lib.PeerExchange(self.term_ptrs.data_ptr(), self.flag_ptrs.data_ptr(), self.rank, self.world, self.parity)

    def descriptor(self):
        """The exchange as the ten integers `torch.ops.cldet.focal_loss` takes (csrc/cldet_torch.cpp, `peer`), for ONE step:
        advances the parity and that parity's use count like exchange() + wait() do."""
        self.check()
        parity = self.parity
        self.uses[parity] += 1
        self.parity ^= 1
        return [self.term_ptrs.data_ptr(), self.flag_ptrs.data_ptr(), self.rank, self.world, parity,
                self._own + 4 * self._flag_off, self._own, (self.uses[parity] * self.n) & 0xFFFFFFFF, self.timeout_ms,
                self.status.data_ptr()]

    def check(self):
        """Raise if an earlier wait on this rank timed out (reads mapped host memory: no device synchronisation)."""
        if int(self.status[0]) != 0:
            raise self._lib.CldetError('peer exchange: a rank did not deliver its loss terms within %d ms (its rows of that step '
                                       'were filled with NaN); the job is out of step or a peer died' % self.timeout_ms)

    def wait(self, stream, out=None):
        """Block `stream` until this parity's gather is complete, copy it into a PRIVATE [4, world*N] tensor (rows bg, fg, reg,
        enhance in global image order) and advance the parity."""
        self.check()
        if out is None:
            out = torch.empty((4, self.world * self.n), dtype=torch.float32, device=self.device)
        self.uses[self.parity] += 1
        target = (self.uses[self.parity] * self.n) & 0xFFFFFFFF
        self._lib.check(self._lib.load().cldet_peer_wait(self._own + 4 * self._flag_off, self._own, self.world, self.n, self.parity,
                                                        target, self.timeout_ms, out.data_ptr(), None, self.status.data_ptr(), stream))
        self.parity ^= 1
        return out

    def close(self):
        """Unmap the peers' buffers and free this rank's (call on every rank, after a barrier)."""
        lib = self._lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.cldet_peer_close(p)
            self._opened = []
            if self._own:
                lib.cldet_peer_free(self._own)
                self._own = 0


class _GatherTerms(torch.autograd.Function):
    """local [K, n_r] -> global [K, N] in rank order.  Backward hands each rank the slice of the incoming gradient
    that belongs to its own images: the loss every rank forms from the gathered terms is the same function, so
    d(loss)/d(local terms) is that slice and no reduction is needed."""

    @staticmethod
    def forward(ctx, local, sizes, rank, group):
        k = local.shape[0]
        nmax = max(sizes)
        world = len(sizes)
        ctx.sizes, ctx.rank = sizes, rank
        if world > 1 and min(sizes) == nmax:
            # equal shards (the usual case): no padding, no per-rank slicing -- one collective and one strided copy
            flat = local.new_empty((world * k, nmax))
            dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
            return flat.view(world, k, nmax).permute(1, 0, 2).reshape(k, world * nmax)
        padded = local.new_zeros((k, nmax))
        padded[:, :local.shape[1]] = local
        if world > 1:
            flat = local.new_empty((world * k, nmax))
            dist.all_gather_into_tensor(flat, padded.contiguous(), group=group)
            out = flat.view(world, k, nmax)
        else:
            out = padded.unsqueeze(0)
        return torch.cat([out[r, :, :sizes[r]] for r in range(world)], dim=1)

    @staticmethod
    def backward(ctx, grad):
        start = sum(ctx.sizes[:ctx.rank])
        return grad[:, start:start + ctx.sizes[ctx.rank]].contiguous(), None, None, None


def gather_terms(local, sizes, rank, group=None):
    return _GatherTerms.apply(local, sizes, rank, group)


class ShardedFocalLoss(nn.Module):
    """FocalLoss over an image-sharded batch.  Each rank passes ITS shard (classifications[n_r,A,C], ...); the result
    dict has the reference's layout for the GLOBAL batch: 'cls_loss' = (bg[N], fg[N]) in global image order,
    'reg_loss' = [1] mean over all N images, 'enhance_on_new_loss' = global sum.  'bg_masks' stays local (it is
    consumed against local tensors by the distillation terms).

    The gradients that flow back are d(global loss)/d(local head outputs); SUM parameter gradients over ranks to get
    the single-process gradient (with DDP's averaging, scale the loss by world_size).

    shard_sizes: how a rank learns the other ranks' image counts WITHOUT a per-step collective and host sync:
      'equal'  (default) every rank holds as many images as this one (DistributedSampler semantics) -- no communication;
      a list   the per-rank image counts, fixed;
      'gather' all-gather the counts on every call (one small collective + a host sync per step; ragged, changing shards).
    """

    def __init__(self, local_loss=None, group=None, use_peer_memory=True, shard_sizes='equal'):
        super().__init__()
        if local_loss is None:
            from .losses import FocalLoss
            local_loss = FocalLoss(upstream_hint='mean')
        self.local_loss = local_loss
        self.group = group
        self.use_peer_memory = use_peer_memory
        if not (shard_sizes in ('equal', 'gather') or isinstance(shard_sizes, (list, tuple))):
            raise ValueError("shard_sizes must be 'equal', 'gather' or a list of per-rank image counts")
        self.shard_sizes = shard_sizes
        self._peer = {}          # n_local -> PeerGather, or False when the mapping could not be set up
        self._weights = {}       # (n_local, n_global, device index) -> [4, n_local] upstream weights baked by the kernel

    def _sizes(self, n_local, world, device):
        if isinstance(self.shard_sizes, (list, tuple)):
            if len(self.shard_sizes) != world:
                raise ValueError('shard_sizes has %d entries for %d ranks' % (len(self.shard_sizes), world))
            return [int(x) for x in self.shard_sizes]
        if self.shard_sizes == 'equal' or world == 1:
            return [n_local] * world
        t = torch.tensor([n_local], dtype=torch.int64, device=device)
        all_n = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_n, t, group=self.group)
        return [int(x.item()) for x in all_n]

    def _hint(self, n_local, n_global, device):
        key = (n_local, n_global, device.index)
        w = self._weights.get(key)
        if w is None:
            w = torch.full((4, n_local), 1.0 / n_global, dtype=torch.float32)
            w[3] = 1.0
            w = self._weights[key] = w.to(device)
        return w

    def _peer_for(self, n_local, sizes, device):
        """Fused all-gather over peer memory needs equal shards, CUDA tensors and the built-in FocalLoss."""
        from .losses import FocalLoss
        if not self.use_peer_memory or device.type != 'cuda' or not isinstance(self.local_loss, FocalLoss):
            return None
        if len(set(sizes)) != 1:
            return None
        if n_local not in self._peer:
            try:
                self._peer[n_local] = PeerGather(n_local, device, self.group)     # collective; agrees on ok / not ok itself
            except Exception:     # no IPC between the ranks (different nodes / containers): use the NCCL all-gather
                self._peer[n_local] = False
        return self._peer[n_local] or None

    def forward(self, classifications, regressions, anchors, annotations, cur_state, params, progress=-1):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n_local = classifications.shape[0]
        sizes = self._sizes(n_local, world, classifications.device)
        if sizes[rank] != n_local:
            raise ValueError('this rank holds %d images but shard_sizes says %d' % (n_local, sizes[rank]))
        n_global = sum(sizes)
        peer = self._peer_for(n_local, sizes, classifications.device) if world > 1 else None
        if peer is not None:
            # fused path: the loss kernel scatters the per-image terms to every rank; the result already holds GLOBAL rows
            old = self.local_loss.upstream_hint
            self.local_loss.upstream_hint = self._hint(n_local, n_global, classifications.device)
            try:
                out = self.local_loss(classifications, regressions, anchors, annotations, cur_state, params, progress, peer=peer)
            finally:
                self.local_loss.upstream_hint = old
            return out       # rows are already global: FocalLoss formed reg_loss / enhance over all N images
        if hasattr(self.local_loss, 'upstream_hint') and isinstance(self.local_loss.upstream_hint, str):
            # the caller's mean runs over the GLOBAL batch: bake 1/N_global into the fused gradients
            self.local_loss.upstream_hint = self._hint(n_local, n_global, classifications.device)
            try:
                out = self.local_loss(classifications, regressions, anchors, annotations, cur_state, params, progress)
            finally:
                self.local_loss.upstream_hint = 'mean'
        else:
            out = self.local_loss(classifications, regressions, anchors, annotations, cur_state, params, progress)
        bg, fg = out['cls_loss']
        # per-image regression terms: the reference returns only their mean; recover reg_j * 1 from the local mean
        reg_local = getattr(self.local_loss, 'last_reg_per_image', None)
        if reg_local is None:
            reg_local = out['reg_loss'].expand(n_local)          # generic local loss: treat the mean as every image's term
        terms = [bg, fg, reg_local]
        has_enh = 'enhance_on_new_loss' in out
        if has_enh:
            terms.append(out['enhance_on_new_loss'].reshape(1).expand(n_local) / max(n_local, 1))
        g = gather_terms(torch.stack(terms), sizes, rank, self.group)
        result = {'cls_loss': (g[0], g[1]), 'reg_loss': g[2].mean(dim=0, keepdim=True)}
        if has_enh:
            result['enhance_on_new_loss'] = g[3].sum()
        if 'bg_masks' in out:
            result['bg_masks'] = out['bg_masks']
        return result

    def close(self):
        """Release the peer-exchange buffers (collective-free; call on every rank after a barrier)."""
        for p in self._peer.values():
            if p:
                p.close()
        self._peer = {}
