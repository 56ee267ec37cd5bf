"""N > 1 host logic on CPU: world_size-2 gloo processes shard channels (and time blocks) exactly like
bench.py does under torchrun, run the ORACLE chain on their share, and the gathered result must equal
the single-process stream.  No GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quisk_b200 import shard


def test_channel_range_partitions():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1024, 1025, 4099):
            got = [shard.channel_range(r, world, n) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            for a, b in zip(got, got[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1


def test_cascade_halo_matches_survey():
    assert shard.cascade_halo([(45, 2)] * 4 + [(98, 2)]) == 2212


def test_time_blocks_cover_stream():
    blocks = shard.time_blocks(153600 + 17, 8, 32, 2212)
    assert blocks[0].start == 0 and blocks[-1].stop == 153600 + 17
    assert sum(b.out_count for b in blocks) == 153600 // 32
    for a, b in zip(blocks, blocks[1:]):
        assert a.stop == b.start and b.start % 32 == 0 and b.out_start == a.out_start + a.out_count
        assert b.halo_start == b.start - 2240 and b.halo_start % 32 == 0      # 2212 rounded up to the decimation


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_channels, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import quisk_oracle as O
    from tests.util import golden
    tabs = golden("quisk_tables.npz")
    # --- channel sharding: every rank decimates its own receivers, no exchange on the data path
    lo, hi = shard.channel_range(rank, world, n_channels)
    outs = []
    for c in range(lo, hi):
        dec = O.ProcessDecimate(192000, tabs)
        outs.append(dec(O.synth_iq(4096, 300 + c, 1.0)))
    mine = np.stack(outs)
    sizes = [None] * world
    dist.all_gather_object(sizes, (lo, hi))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # result collection only (bench gathers nothing)
    # --- time-block sharding of ONE stream with halos: rank computes its outputs from halo_start
    x = O.synth_iq(32768, 999, 1.0)
    blk = shard.time_blocks(len(x), world, 4, shard.cascade_halo([(45, 2), (98, 2)]))[rank]
    dec = O.ProcessDecimate(192000, tabs)           # 192 k -> HB45 -> FIR/2 -> 48 k (total decimation 4)
    y = dec(x[blk.halo_start:blk.stop])
    y = y[len(y) - blk.out_count:] if blk.out_count else y[:0]
    parts = [None] * world
    dist.all_gather_object(parts, (blk.out_start, y))
    # --- the C5 channelizer sharded the same way: a rank starts at halo_start with the ABSOLUTE sample index (receiver
    #     phases and frame alignment follow from it, quisk_cuda_pfb_seek) and drops the frames inside its halo
    K, D, P = 64, 32, 4
    proto = np.hanning(K * P) / (K * P)
    xw = O.synth_iq(64 * D, 998, 1.0)
    tb = shard.time_blocks(len(xw), world, D, K * P - 1)[rank]
    yc = O.channelizer_oracle(xw[tb.halo_start:tb.stop], [0, 5, K - 1], proto, K, D, n0=tb.halo_start)
    yc = yc[:, yc.shape[1] - tb.out_count:]
    cparts = [None] * world
    dist.all_gather_object(cparts, (tb.out_start, yc))
    # max-over-ranks reduction of a per-rank time, as bench.py does
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((np.concatenate(gathered), sizes, parts, float(t.item()), cparts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_sharding_matches_single_process():
    from oracle import quisk_oracle as O
    from tests.util import golden
    tabs = golden("quisk_tables.npz")
    world, n_channels = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, q)) for r in range(world)]
    for p in procs: p.start()
    got, sizes, parts, tmax, cparts = q.get(timeout=60)
    for p in procs: p.join(timeout=30)
    assert all(p.exitcode == 0 for p in procs)
    assert sizes == [(0, 3), (3, 5)] and tmax == 2.0
    ref = np.stack([O.ProcessDecimate(192000, tabs)(O.synth_iq(4096, 300 + c, 1.0)) for c in range(n_channels)])
    assert np.array_equal(got, ref)
    # time-block sharding: stitched outputs equal the sequential stream (same arithmetic on the same inputs)
    x = O.synth_iq(32768, 999, 1.0)
    seq = O.ProcessDecimate(192000, tabs)(x)
    stitched = np.concatenate([y for _, y in sorted(parts, key=lambda t: t[0])])
    assert len(stitched) == len(seq) == 8192
    assert O.rel_rms(stitched, seq) < 1e-14
    # channelizer: the stitched shards are the sequential frames exactly (same arithmetic on the same samples)
    K, D, P = 64, 32, 4
    proto = np.hanning(K * P) / (K * P)
    xw = O.synth_iq(64 * D, 998, 1.0)
    cseq = O.channelizer_oracle(xw, [0, 5, K - 1], proto, K, D)
    cst = np.concatenate([y for _, y in sorted(cparts, key=lambda t: t[0])], axis=1)
    assert cst.shape == cseq.shape == (3, 64)
    assert np.array_equal(cst, cseq)
