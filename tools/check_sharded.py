#!/usr/bin/env python
"""Multi-GPU check of the image-sharded loss (run under torchrun, one rank per GPU, NCCL):
every rank computes its shard with ShardedFocalLoss; rank 0 also computes the whole batch on one GPU; the caller-side
reductions (mean and the clip_loss mask) and the gradients of each shard must agree.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cl_object_detection_b200 as cld  # noqa: E402
from bench import synth_annotations  # noqa: E402


def caller_reduction(out, clip):
    bg, fg = out['cls_loss']
    mask = fg >= clip
    fg_term = fg[mask].mean() if mask.sum() > 0 else fg.sum() * 0
    return bg.mean() + fg_term + out['reg_loss'].mean()


def straggler(sharded, p, r, anchors, ann, params, rank, world):
    """A rank that is late by more than the exchange's timeout: the waiting ranks must see NaN in the late rank's rows of THAT
    step and an exception on their next call -- never stale numbers -- and the late rank itself completes normally."""
    import time
    pg = [v for v in sharded._peer.values() if v][0]
    pg.timeout_ms = 300
    dist.barrier()
    torch.cuda.synchronize()
    if rank == world - 1:
        time.sleep(1.5)
    with torch.no_grad():
        out = sharded(p, r, anchors, ann, 0, params)
    torch.cuda.synchronize()
    bg = out['cls_loss'][0]
    n = p.shape[0]
    late = bg[(world - 1) * n:]
    if rank != world - 1:
        assert bool(torch.isnan(late).all()), 'a late rank\'s rows must be poisoned, got %r' % (late,)
        assert not bool(torch.isnan(bg[:(world - 1) * n]).any())
        try:
            sharded(p, r, anchors, ann, 0, params)
            raise AssertionError('the step after a timed-out exchange must raise')
        except cld.CldetError as e:
            print('rank %d: straggler detected as designed: %s' % (rank, str(e)[:90]))
    else:
        assert not bool(torch.isnan(bg).any()), 'the late rank itself received everything'
        print('rank %d: late rank completed normally' % rank)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    equal = os.environ.get('CLDET_EQUAL_SHARDS', '0') == '1'        # equal shards take the fused peer-memory all-gather
    h, w, c, g = 512, 512, 20, 12
    n_global = 4 * world if equal else 5 * world + 1               # uneven shards on purpose (NCCL all-gather path)
    anchors = cld.generate_anchors(h, w, dev)
    a = anchors.shape[1]
    gen = torch.Generator(device='cpu').manual_seed(77)
    probs = torch.sigmoid(torch.randn(n_global, a, c, generator=gen) * 2 - 4)
    reg = torch.randn(n_global, a, 4, generator=gen)
    ann = torch.from_numpy(synth_annotations(np.random.default_rng(77), n_global, g, h, w, c, empty=(1,)))
    params = cld.HeadParams()
    sl = cld.shard_slice(n_global, world, rank)
    p = probs[sl].to(dev).requires_grad_(True)
    r = reg[sl].to(dev).requires_grad_(True)
    sharded = cld.ShardedFocalLoss() if equal else cld.ShardedFocalLoss(shard_sizes=cld.shard_sizes(n_global, world))
    out = sharded(p, r, anchors, ann[sl].to(dev), 0, params)
    used_peer = bool(sharded._peer and any(v for v in sharded._peer.values()))
    if equal:      # run three times more: both parities of the gather buffer, several uses of each (growing arrival targets)
        first = [t.clone() for t in out['cls_loss']]
        for _ in range(3):
            p.grad = None
            r.grad = None
            out = sharded(p, r, anchors, ann[sl].to(dev), 0, params)
        # results of an earlier step stay what they were: they are private copies, not views of the exchange buffer
        assert all(torch.equal(a, b) for a, b in zip(first, [t for t in out['cls_loss']])), 'step results are not reproducible'
    clip = 0.5 * float(out['cls_loss'][1].detach().median())
    loss = caller_reduction(out, clip)
    loss.backward()
    # single-GPU truth on every rank (cheap at this size)
    p0 = probs.to(dev).requires_grad_(True)
    r0 = reg.to(dev).requires_grad_(True)
    ref_out = cld.FocalLoss()(p0, r0, anchors, ann.to(dev), 0, params)
    ref = caller_reduction(ref_out, clip)
    ref.backward()
    # per-image terms: same kernels, but the block partition (hence the fp32 summation order) depends on the local batch size
    assert torch.allclose(out['cls_loss'][0], ref_out['cls_loss'][0], rtol=2e-6, atol=0), 'bg terms differ'
    assert torch.allclose(out['cls_loss'][1], ref_out['cls_loss'][1], rtol=2e-6, atol=0), 'fg terms differ'
    assert torch.allclose(loss, ref, rtol=1e-6), (loss, ref)
    assert torch.allclose(p.grad, p0.grad[sl], rtol=1e-6, atol=1e-12), 'cls gradient differs'
    assert torch.allclose(r.grad, r0.grad[sl], rtol=1e-6, atol=1e-12), 'reg gradient differs'
    if equal and used_peer and os.environ.get('CLDET_STRAGGLER', '0') == '1':
        straggler(sharded, p, r, anchors, ann[sl].to(dev), params, rank, world)
    dist.barrier()
    if rank == 0:
        print('sharded loss ok: world=%d n_global=%d loss=%.6f fused_peer_allgather=%s' % (world, n_global, float(loss), used_peer))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
