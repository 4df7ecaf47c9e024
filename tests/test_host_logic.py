"""Host-side logic that needs no GPU: checkpoint compatibility of the RAFT module, padding, the
pair sharding (incl. a world_size-2 gloo run of the optional gather)."""
import json
import os

import numpy as np
import pytest
import torch

from sd_animation_optical_flow_b200 import shard
from sd_animation_optical_flow_b200.raft import RAFT, InputPadder, convex_upsample, coords_grid, fill_weights_by_name

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('kind', ['basic', 'small'])
def test_state_dict_keys_match_reference(kind):
    ref = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'raft_state_dict_keys.json')))[kind]
    from types import SimpleNamespace
    sd = RAFT(SimpleNamespace(small=(kind == 'small'))).state_dict()
    assert sorted(sd) == sorted(ref)
    for k, v in sd.items():
        assert list(v.shape) == ref[k], k


def test_dataparallel_checkpoint_prefix_is_accepted():
    m = RAFT()
    sd = {'module.' + k: v for k, v in m.state_dict().items()}
    RAFT().load_state_dict(sd)   # ofgen.py:67-68 loads a DataParallel checkpoint


def test_fill_weights_is_order_independent():
    a, b = RAFT(), RAFT()
    fill_weights_by_name(a, 3)
    fill_weights_by_name(b, 3)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k


def test_input_padder():
    p = InputPadder((1, 3, 132, 150))
    x = torch.arange(132 * 150, dtype=torch.float32).reshape(1, 1, 132, 150)
    y, = p.pad(x)
    assert y.shape[-2:] == (136, 152)
    assert torch.equal(p.unpad(y), x)
    assert InputPadder((1, 3, 768, 512)).pad(torch.zeros(1, 3, 768, 512))[0].shape[-2:] == (768, 512)


def test_coords_grid_is_xy():
    g = coords_grid(1, 3, 5, 'cpu')
    assert g.shape == (1, 2, 3, 5)
    assert g[0, 0, 2, 4] == 4 and g[0, 1, 2, 4] == 2


def test_convex_upsample_of_constant_flow():
    flow = torch.ones(1, 2, 4, 5) * torch.tensor([1.5, -2.0]).view(1, 2, 1, 1)
    mask = torch.randn(1, 576, 4, 5)
    up = convex_upsample(flow, mask)
    assert up.shape == (1, 2, 32, 40)
    # interior: convex combination of a constant = 8 * constant
    assert torch.allclose(up[0, :, 8:-8, 8:-8], (8 * flow[0, :, :1, :1]).expand(2, 16, 24), atol=1e-5)


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 31, 256, 999):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_shard_by_target_keeps_references_together():
    pairs = [(s, t) for t in range(10) for s in (0, 25, 50) if s != t]
    seen = []
    for r in range(4):
        mine = shard.shard_by_target(pairs, r, 4)
        targets = {t for _, t in mine}
        for t in targets:
            assert sum(1 for p in mine if p[1] == t) == sum(1 for p in pairs if p[1] == t)
        seen += mine
    assert sorted(seen) == sorted(pairs)


def test_key_frame_pairs():
    pairs = shard.key_frame_pairs(1000, 25)
    assert len(pairs) == 960 and pairs[0] == (0, 1) and pairs[-1] == (975, 999)


def _gather_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    s, e = shard.shard_range(n_total, rank, world)
    local = torch.arange(s, e, dtype=torch.float32)[:, None, None].expand(e - s, 2, 3).contiguous()
    out = shard.gather_stack(local, n_total)
    q.put((rank, out[:, 0, 0].tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [5, 8])
def test_gather_stack_world2_gloo(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r] == [float(i) for i in range(n_total)]
