"""The N>1 plumbing on CPU: world_size-2 gloo, interleaved shards + the single all-gather of result records."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from ccp_b200 import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rb = 2656
    idx = sweep.local_indices(n_total, rank, world)
    # fake "results": record f carries f in its zero_count field and f*3 in failure_count
    local = np.zeros((len(idx), rb), dtype=np.uint8)
    hdr = local[:, :12].view(np.int32)
    hdr[:, 1] = idx
    hdr[:, 2] = [3 * i for i in idx]
    full = sweep.gather_records(torch.from_numpy(local.reshape(-1)), n_total, rank, world, rb)
    st, zc, fc = sweep.counts_from_records(full)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.stack([st, zc, fc]))
    # the copy-free form bench.py times: preallocated output, the collective's own [world, m_pad, record] layout, front order on the host
    m_pad = sweep.padded_shard_len(n_total, world)
    pre = torch.empty(world * m_pad * rb, dtype=torch.uint8)
    raw = sweep.gather_records(torch.from_numpy(local.reshape(-1)), n_total, rank, world, rb, out=pre, ordered=False)
    assert raw.shape == (world, m_pad, rb) and raw.data_ptr() == pre.data_ptr()
    assert np.array_equal(sweep.front_order(raw, n_total).reshape(-1), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_restores_front_order(tmp_path):
    world, n_total = 2, 101          # ragged: ranks hold 51 and 50 records
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        a = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert a.shape == (3, n_total)
        assert np.array_equal(a[1], np.arange(n_total))
        assert np.array_equal(a[2], 3 * np.arange(n_total))
        assert np.all(a[0] == 0)


def test_single_rank_gather_is_identity():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_package()
    from ccp_b200 import sweep
    local = torch.arange(5 * 2656, dtype=torch.int64).to(torch.uint8)
    out = sweep.gather_records(local, 5, 0, 1, 2656)
    assert torch.equal(out, local)
