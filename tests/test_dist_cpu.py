"""world_size-2 gloo tests (CPU) of the image-sharded loss wrapper's host logic: shard rule, gather order, backward
slicing, and that the caller-side reductions (mean / clip_loss mask, losses.py:575-588) over the gathered terms equal
the single-process result.  The local per-image loss is a differentiable stand-in (the CUDA kernels need a GPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from cl_object_detection_b200.dist import ShardedFocalLoss, gather_terms, shard_sizes, shard_slice


class StandInLoss(nn.Module):
    """Per-image terms that depend only on that image's tensors, same result layout as FocalLoss."""
    upstream_hint = 'mean'

    def forward(self, cls, reg, anchors, ann, cur_state, params, progress=-1):
        bg = (cls ** 2).sum(dim=(1, 2))
        fg = cls.abs().sum(dim=(1, 2)) * 0.01
        reg_j = (reg ** 2).mean(dim=(1, 2))
        self.last_reg_per_image = reg_j
        return {'cls_loss': (bg, fg), 'reg_loss': reg_j.mean(dim=0, keepdim=True)}


def caller_reduction(out, clip):
    bg, fg = out['cls_loss']
    mask = fg >= clip
    fg_term = fg[mask].mean() if mask.sum() > 0 else fg.sum() * 0
    return bg.mean() + fg_term + out['reg_loss'].mean()


def _worker(rank, world, port, n_global, tmp, mode):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cls = torch.rand(n_global, 7, 3, generator=g)
        reg = torch.randn(n_global, 7, 4, generator=g)
        clip = float(cls.abs().sum(dim=(1, 2)).median() * 0.01)
        # single-process truth
        c0 = cls.clone().requires_grad_(True)
        r0 = reg.clone().requires_grad_(True)
        ref = caller_reduction(StandInLoss()(c0, r0, None, None, 0, None), clip)
        ref.backward()
        # sharded
        sl = shard_slice(n_global, world, rank)
        c = cls[sl].clone().requires_grad_(True)
        r = reg[sl].clone().requires_grad_(True)
        sizes = {'equal': 'equal', 'gather': 'gather', 'list': shard_sizes(n_global, world)}[mode]
        sharded = ShardedFocalLoss(StandInLoss(), shard_sizes=sizes)
        out = sharded(c, r, None, None, 0, None)
        assert out['cls_loss'][0].shape[0] == n_global
        loss = caller_reduction(out, clip)
        loss.backward()
        assert torch.allclose(loss, ref, rtol=1e-6)
        assert torch.allclose(c.grad, c0.grad[sl], rtol=1e-6, atol=1e-9)
        assert torch.allclose(r.grad, r0.grad[sl], rtol=1e-6, atol=1e-9)
        # gather order = global image order
        ids = torch.arange(n_global, dtype=torch.float32)[sl].reshape(1, -1)
        got = gather_terms(ids, shard_sizes(n_global, world), rank)
        assert torch.equal(got[0], torch.arange(n_global, dtype=torch.float32))
        open(os.path.join(tmp, 'ok%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_global,mode', [(4, 'equal'), (5, 'list'), (5, 'gather')])
def test_sharded_loss_world2_gloo(tmp_path, n_global, mode):
    # 'equal' (the default) assumes DistributedSampler-style equal shards and needs no per-step collective; ragged shards
    # are given as a list or gathered on every call
    port = 29500 + (os.getpid() % 2000) + n_global + (7 if mode == 'gather' else 0)
    mp.spawn(_worker, args=(2, port, n_global, str(tmp_path), mode), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'ok0') and os.path.exists(tmp_path / 'ok1')


def test_shard_rule():
    assert shard_sizes(16, 8) == [2] * 8
    assert shard_sizes(5, 2) == [3, 2]
    assert shard_slice(5, 2, 1) == slice(3, 5)
    assert sum(shard_sizes(33, 8)) == 33


def test_single_process_passthrough():
    cls = torch.rand(3, 5, 2, requires_grad=True)
    reg = torch.randn(3, 5, 4, requires_grad=True)
    a = ShardedFocalLoss(StandInLoss())(cls, reg, None, None, 0, None)
    b = StandInLoss()(cls, reg, None, None, 0, None)
    assert torch.allclose(a['cls_loss'][0], b['cls_loss'][0]) and torch.allclose(a['reg_loss'], b['reg_loss'])


def test_shard_sizes_argument_is_checked():
    with pytest.raises(ValueError):
        ShardedFocalLoss(StandInLoss(), shard_sizes='sometimes')
    cls = torch.rand(3, 5, 2)
    reg = torch.randn(3, 5, 4)
    with pytest.raises(ValueError):          # this rank holds 3 images, the list says 2
        ShardedFocalLoss(StandInLoss(), shard_sizes=[2])(cls, reg, None, None, 0, None)
