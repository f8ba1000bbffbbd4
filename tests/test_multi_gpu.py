"""N > 1: the row-shard plumbing on CPU (gloo, world_size 2) and, when two GPUs are visible, the
sharded training step against the single-GPU one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sslrec_b200.parallel import RowShard
    sh = RowShard(dist, rank, world, n)
    full = torch.arange(n * 2 * 3, dtype=torch.float32).view(n, 2, 3)
    local = torch.zeros(sh.block, 2, 3)
    local[:sh.n_local] = full[sh.r0:sh.r1]
    got = sh.allgather_rows(local)
    assert torch.equal(got, full)
    # every global row belongs to exactly one rank; side ranges split without gaps
    owned = torch.zeros(n)
    owned[sh.r0:sh.r1] = 1
    dist.all_reduce(owned)
    assert torch.equal(owned, torch.ones(n))
    n_user = n // 3
    lo_u, hi_u = sh.side_range(0, n_user)
    lo_i, hi_i = sh.side_range(n_user, n - n_user)
    cnt = torch.tensor([hi_u - lo_u, hi_i - lo_i], dtype=torch.float32)
    dist.all_reduce(cnt)
    assert cnt.tolist() == [n_user, n - n_user]
    # partial (row sum, weighted sum) of a sharded softmax contraction all-reduce to the full one
    g = torch.Generator().manual_seed(0)
    a, t = torch.randn(5, 4, generator=g), torch.randn(n, 4, generator=g)
    lo, hi = sh.side_range(0, n)
    e = torch.exp(a @ t[lo:hi].T)
    red = torch.cat([e @ t[lo:hi], e.sum(1, keepdim=True)], 1)
    blk = torch.zeros(sh.side_block(n), 4)
    blk[:hi - lo] = t[lo:hi]
    assert torch.equal(sh.allgather_side(blk, n), t)
    sh.allreduce_sum(red)
    ef = torch.exp(a @ t.T)
    assert torch.allclose(red[:, :4], ef @ t, rtol=1e-5) and torch.allclose(red[:, 4], ef.sum(1), rtol=1e-5)
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [10, 11, 64])
def test_row_shard_plumbing_gloo_world2(n):
    mp.spawn(_gloo_worker, args=(2, _free_port(), n), nprocs=2, join=True)


def _dp_gloo_worker(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sslrec_b200.parallel import BatchShard
    sync = BatchShard(dist, rank, world)
    # two parameters whose gradients are the halves of one flat sink (the embedding table) + a separate one
    flat = torch.arange(12, dtype=torch.float32).view(6, 2) * (rank + 1)
    pu, pi, pw = (torch.nn.Parameter(torch.zeros(4, 2)), torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(3)))
    pu.grad, pi.grad, pw.grad = flat[:4], flat[4:], torch.full((3,), float(rank))
    bufs = sync.coalesce([pi.grad, pw.grad, pu.grad])
    assert sorted(b.numel() for b in bufs) == [3, 12]            # the adjacent halves travel as one buffer
    sync.average_gradients([pu, pi, pw, torch.nn.Parameter(torch.zeros(1))])   # a parameter without .grad is skipped
    want = torch.arange(12, dtype=torch.float32).view(6, 2) * (sum(range(1, world + 1)) / world)
    assert torch.equal(flat, want) and torch.equal(pu.grad, want[:4]) and torch.equal(pi.grad, want[4:])
    assert torch.allclose(pw.grad, torch.full((3,), (world - 1) / 2.0))
    # the loader gives every rank a disjoint share of each epoch
    ds = torch.utils.data.TensorDataset(torch.arange(20))
    loader = sync.shard_loader(ds, batch_size=4, seed=7)
    loader.sampler.set_epoch(3)
    mine = torch.cat([b[0] for b in loader])
    seen = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(seen, mine)
    assert sorted(torch.cat(seen).tolist()) == list(range(20))
    dist.destroy_process_group()


def test_batch_shard_gradient_average_gloo_world2():
    mp.spawn(_dp_gloo_worker, args=(2, _free_port()), nprocs=2, join=True)


def _dp_gpu_worker(rank, world, port, out):
    """Data-parallel step (one batch per rank, averaged gradients) == the single-GPU step on the concatenated batch."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    import ssl_test_helpers as H
    from oracle import inputs, replay
    from sslrec_b200.optim import FusedAdam
    from sslrec_b200.parallel import BatchShard
    for name in ('simgcl', 'lightgcn'):
        g = replay.load_golden(name, 'small')
        case = inputs.make_case('small')
        full = [torch.from_numpy(case[k]).cuda() for k in ('ancs', 'poss', 'negs')]
        half = full[0].numel() // world
        out = {}
        for dp in (False, True):
            model, _ = H.make_model(name, case, g['hp'], device=f'cuda:{rank}')
            model.load_state_dict({'user_embeds': case['user_e'], 'item_embeds': case['item_e']})
            opt = FusedAdam(model.parameters(), lr=1e-2)
            batch = [t[rank * half:(rank + 1) * half] for t in full] if dp else [t[:world * half] for t in full]
            loss, _ = model.cal_loss(batch)
            loss.backward()
            if dp:
                BatchShard(dist, rank, world).average_gradients(model.parameters())
                dist.all_reduce(loss, op=dist.ReduceOp.AVG)
            grads = (model.user_embeds.grad.clone(), model.item_embeds.grad.clone())
            opt.step()
            out[dp] = (loss.item(), grads, model.user_embeds.detach().clone(), model.item_embeds.detach().clone())
        assert abs(out[True][0] - out[False][0]) <= 2e-6 * max(1.0, abs(out[False][0])), name
        for a, b in zip(out[True][1], out[False][1]):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-7 + 2e-5 * b.abs().max().item()), name
        # identical parameters on every rank after the step
        for p in out[True][2:]:
            ref = p.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(p, ref), name
    dist.destroy_process_group()


@pytest.mark.gpu
def test_data_parallel_step_matches_big_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    mp.spawn(_dp_gpu_worker, args=(2, _free_port(), None), nprocs=2, join=True)


def _gpu_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    import ssl_test_helpers as H
    from oracle import inputs, replay
    from sslrec_b200.parallel import RowShard
    g = replay.load_golden('simgcl', 'small')
    case = inputs.make_case('small')
    res = {}
    for sharded in (False, 'loss', 'all'):
        model, _ = H.make_model('simgcl', case, g['hp'], device=f'cuda:{rank}')
        model.load_state_dict({'user_embeds': case['user_e'], 'item_embeds': case['item_e']})
        if sharded:
            model.comm = RowShard(dist, rank, world, case['n_user'] + case['n_item'], shard_propagation=(sharded == 'all'))
            model._plans.clear()
        batch = [torch.from_numpy(case[k]).cuda() for k in ('ancs', 'poss', 'negs')]
        loss, _ = model.cal_loss(batch)
        loss.backward()
        res[sharded] = (loss.item(), model.user_embeds.grad.clone(), model.item_embeds.grad.clone())
    for mode in ('loss', 'all'):
        assert abs(res[mode][0] - res[False][0]) <= 1e-6 * max(1.0, abs(res[False][0])), mode
        for a, b in zip(res[mode][1:], res[False][1:]):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-7 + 1e-5 * b.abs().max().item()), mode
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_step_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    mp.spawn(_gpu_worker, args=(2, _free_port(), None), nprocs=2, join=True)
