"""Single-GPU tests of the peer-memory kernels of the row-sharded path (csrc/t4r_peer.cu): the "ranks" are
separate allocations of ONE process (``ops.LocalPeerView``), so the arithmetic of the sharded lookup, the label-row pull
and the cross-shard softmax combine is pinned on the driver's one-GPU box; the multi-process choreography (CUDA IPC,
the two ordering collectives) is tests/test_gpu_sharded.py.  Integer / copy results bit-exact, losses within 1e-4 of the
CPU oracle (the bar of the replicated head tests), ranks exact through Recall@k to 1e-6."""
import pytest
import torch

import t4r_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from transformers4rec_b200 import ops as _ops
    return _ops


def _shards(table, world):
    per = (table.shape[0] + world - 1) // world
    return per, [table[r * per:(r + 1) * per].contiguous().cuda() for r in range(world)]


@pytest.mark.parametrize("V,K,world", [(1000, 64, 1), (1003, 256, 3), (7919, 200, 8), (50, 4, 16), (4001, 512, 4)])
def test_peer_gather_rows_bit_exact(ops, V, K, world):
    g = torch.Generator().manual_seed(V + K)
    table = torch.randn(V, K, generator=g)
    per, shards = _shards(table, world)
    view = ops.LocalPeerView([s if s.numel() else torch.zeros(1, K, device="cuda") for s in shards])
    n = 3000
    ids = torch.randint(0, V, (n,), generator=g)
    ids[::3] = 0                      # the padding id: served from the per-CTA staged row
    ids[5], ids[6] = V - 1, 1
    rows, planes = ops.peer_gather_rows(view, V, per, K, ids.cuda(), pad_id=0)
    ref = table[ids]
    assert torch.equal(rows.cpu(), ref)
    assert torch.equal(planes.cpu(), ops.split_planes(ref.cuda()).cpu())
    # without the padding hint the same rows come out
    rows2, _ = ops.peer_gather_rows(view, V, per, K, ids.cuda(), pad_id=-1, want_planes=False)
    assert torch.equal(rows2.cpu(), ref)
    # device-side count: valid rows, then zeros up to the next multiple of 256
    cnt = torch.tensor([1234], dtype=torch.int32, device="cuda")
    rows3, planes3 = ops.peer_gather_rows(view, V, per, K, ids.cuda(), count=cnt, pad_id=0)
    assert torch.equal(rows3[:1234].cpu(), ref[:1234]) and not rows3[1234:1280].any() and not planes3[:, 1234:1280].any()
    # ids outside [0, V): zero rows and the error flag
    bad = ids.clone()
    bad[10], bad[11] = V, -7
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rows4, _ = ops.peer_gather_rows(view, V, per, K, bad.cuda(), pad_id=0, want_planes=False, err_flag=err)
    assert err.item() == 1 and not rows4[10:12].any() and torch.equal(rows4[12:].cpu(), ref[12:])


def test_peer_gather_rows_rejects_unsupported_widths(ops):
    from transformers4rec_b200 import T4RError
    t = torch.zeros(10, 6, device="cuda")
    with pytest.raises(T4RError):
        ops.peer_gather_rows(ops.LocalPeerView([t]), 10, 10, 6, torch.zeros(4, dtype=torch.long, device="cuda"))


@pytest.mark.parametrize("world,cap,K,counts", [(1, 300, 64, [123]), (3, 512, 256, [400, 0, 512]),
                                                 (8, 700, 128, [1, 700, 33, 256, 0, 699, 257, 5])])
def test_peer_pull_rows_bit_exact(ops, world, cap, K, counts):
    g = torch.Generator().manual_seed(world * 1000 + cap)
    xs = [torch.randn(cap, K, generator=g) for _ in range(world)]
    ys = [torch.randint(1, 10**12, (cap,), generator=g) for _ in range(world)]
    mx = ops.LocalPeerView([x.cuda() for x in xs], rank=world - 1)
    my = ops.LocalPeerView([y.cuda() for y in ys], rank=world - 1)
    cdev = torch.tensor(counts, dtype=torch.int32, device="cuda")
    out = ops.peer_pull_rows(mx, my, cdev, cap, K)
    ref_x = torch.cat([xs[r][:counts[r]] for r in range(world)])
    ref_y = torch.cat([ys[r][:counts[r]] for r in range(world)])
    T = sum(counts)
    assert out["t_total"].item() == T and out["my_start"].item() == sum(counts[:-1])
    assert torch.equal(out["x"][:T].cpu(), ref_x) and torch.equal(out["labels"][:T].cpu(), ref_y)
    fill = min(world * cap, (T + 255) // 256 * 256)
    assert not out["x"][T:fill].any() and not out["planes"][:, T:fill].any()
    assert torch.equal(out["planes"][:, :T].cpu(), ops.split_planes(ref_x.cuda()).cpu())


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("nprod", [3, 2])
def test_sharded_head_on_one_gpu_matches_oracle(ops, world, nprod):
    """The row-sharded full-softmax head with every shard computed in this process: per-shard logits + online LSE over
    V/world rows (``v_offset``), statistics into per-shard windows, ``peer_combine_lse`` across them -- loss, row
    losses and cross-shard label ranks against the CPU oracle's full softmax over the whole table."""
    torch.manual_seed(11)
    T, V, De, tau = 333, 5003, 64, 0.7
    xt, W = torch.randn(T, De), torch.randn(V, De) * 0.2
    y = torch.randint(1, V, (T,))
    ref_loss, ref_logits = O.full_softmax_head(xt, y, W, tau)
    ref_rows = torch.nn.functional.cross_entropy(ref_logits, y, reduction="none")
    cap_g = 512
    x_pad = torch.zeros(cap_g, De); x_pad[:T] = xt
    y_pad = torch.zeros(cap_g, dtype=torch.long); y_pad[:T] = y
    xd, yd = x_pad.cuda(), y_pad.cuda()
    t_total = torch.tensor([T], dtype=torch.int32, device="cuda")
    per, shards = _shards(W, world)
    # the labels' logits over the whole table (evaluation): W[y] through the sharded lookup
    wy, _ = ops.peer_gather_rows(ops.LocalPeerView(shards), V, per, De, yd, count=t_total, want_planes=False)
    tgt = ops.label_logit(xd, wy, torch.arange(cap_g, device="cuda"), t_dev=t_total, inv_temperature=1.0 / tau)
    stats = []
    for r, Ws in enumerate(shards):
        st = torch.zeros(3, cap_g, device="cuda")
        if nprod == 2:
            xm, xi = ops.split_planes_mixed(xd)
            wm, wi = ops.split_planes_mixed(Ws)
            ops.head_softmax_ce(xm, xd, yd, wm, Ws, t_dev=t_total, inv_temperature=1.0 / tau, v_offset=r * per,
                                want_loss=False, want_rank=True, rank_tgt=tgt, nprod=2, xt_inv_scale=xi, w_inv_scale=wi,
                                out_stats=st)
        else:
            ops.head_softmax_ce(ops.split_planes(xd), xd, yd, ops.split_planes(Ws), Ws, t_dev=t_total,
                                inv_temperature=1.0 / tau, v_offset=r * per, want_loss=False, want_rank=True,
                                rank_tgt=tgt, out_stats=st)
        stats.append(st)
    row_loss, loss, row_rank = ops.peer_combine_lse(ops.LocalPeerView(stats), cap_g, t_total, with_rank=True)
    assert abs(loss.item() - ref_loss.item()) < 1e-4
    assert (row_loss[:T].cpu() - ref_rows).abs().max().item() < 2e-4 and not row_loss[T:].any()
    ks = [1, 5, 10, 20]
    got = ops.recall_from_ranks(row_rank, ks, t_total).cpu()
    assert (got - O.recall_at_mean(ks, ref_logits, y)).abs().max().item() < 1e-6
