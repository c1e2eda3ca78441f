"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pair dealing, point shards, the partial-row
all-reduce and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from icp_variants_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # one large pair: partial normal-equation rows of the two shards sum to the row of the whole
        rng = np.random.default_rng(5)
        rows = rng.normal(size=(1000, 28))
        sl = parallel.shard_points(len(rows), world, rank)
        total = parallel.allreduce_sum(rows[sl].sum(0))
        ok_sum = np.allclose(total, rows.sum(0), rtol=1e-12, atol=1e-12)
        # pair queue: every pair goes to exactly one rank
        mine = parallel.shard_pairs(44, world, rank)
        counts = parallel.allreduce_sum(np.bincount(mine, minlength=44).astype(np.float64))
        ok_pairs = np.array_equal(counts, np.ones(44))
        # timing reduction
        ok_max = parallel.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        # mailbox handles of the fused (peer-memory) exchange arrive in rank order on every rank
        handles = parallel.gather_peer_handles(bytes([rank]) * 64)
        ok_max = ok_max and handles == [bytes([r]) * 64 for r in range(world)]

        class FakeCtx:                                   # attach_peers drives export -> gather -> attach -> barrier
            def peer_export(self):
                return bytes([100 + rank]) * 64

            def peer_attach(self, r, w, hs):
                self.got = (r, w, list(hs))
        fc = FakeCtx()
        parallel.attach_peers(fc)
        ok_max = ok_max and fc.got == (rank, world, [bytes([100 + r]) * 64 for r in range(world)])
        # dynamic deal: both ranks draw from one ticket counter; every pair is taken exactly once, and a pass with a fresh key
        # starts from zero again (sequence.alignPairs over fake contexts: no GPU needed for the queue's host logic)
        from icp_variants_b200 import sequence, synth

        class QueueCtx:
            def __init__(self):
                self.log, self.busy = [], False

            def set_config(self, c):
                pass

            def set_target(self, p, n, c):
                self.tgt = p

            def set_source(self, p, n, c):
                self.src = p

            def estimate_pose_async(self, pose):
                assert not self.busy                      # a context is drained before it is reused
                self.busy = True

            def estimate_pose_finish(self):
                assert self.busy
                self.busy = False
                pose = np.eye(4, dtype=np.float32)
                pose[0, 3] = float(self.src[0, 0])        # which pair this was
                return pose, 30
        cloud = lambda v: synth.Cloud(np.full((2, 3), v, np.float32), np.zeros((2, 3), np.float32), np.zeros((2, 4), np.uint8))   # noqa: E731
        all_pairs = [(cloud(k), cloud(-k)) for k in range(44)]
        for rep in range(2):
            dist.barrier()
            tickets = parallel.PairTickets(44, key=f"test_pass{rep}")
            res = sequence.alignPairs([QueueCtx() for _ in range(3)], all_pairs, None, tickets=tickets)
            took = np.array([0.0 if r is None else 1.0 for r in res])
            ok_pairs = ok_pairs and all(r is None or (r.nIterations == 30 and r.pose[0, 3] == k) for k, r in enumerate(res))
            ok_pairs = ok_pairs and sorted(tickets.drawn) == [k for k in range(44) if res[k] is not None]
            ok_pairs = ok_pairs and np.array_equal(parallel.allreduce_sum(took), np.ones(44))
        q.put((rank, bool(ok_sum), bool(ok_pairs), bool(ok_max), len(mine)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:4] for r in res] == [(True, True, True)] * world
    assert sum(r[4] for r in res) == 44


def test_shards_cover_everything():
    for n in (0, 1, 7, 44, 370488):
        for world in (1, 2, 4, 8):
            sl = [parallel.shard_points(n, world, r) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            assert max(s.stop - s.start for s in sl) - min(s.stop - s.start for s in sl) <= 1
            idx = sorted(i for r in range(world) for i in range(n)[parallel.shard_points_interleaved(n, world, r)])
            assert idx == list(range(n)) or n > 100000          # (the full check on the big size would take seconds)
            pairs = sorted(i for r in range(world) for i in parallel.shard_pairs(n % 100, world, r))
            assert pairs == list(range(n % 100))


def test_align_pairs_queue_host_logic_without_tickets():
    """The static queue (one process): every pair registered once, in order, contexts drained before they are reused, for fewer,
    as many and more pairs than contexts; a local ticket source gives the same result."""
    from icp_variants_b200 import sequence, synth

    class Ctx:
        def __init__(self):
            self.busy, self.n = False, 0

        def set_config(self, c):
            pass

        def set_target(self, p, n, c):
            pass

        def set_source(self, p, n, c):
            self.tag = float(p[0, 0])

        def estimate_pose_async(self, pose):
            assert not self.busy
            self.busy, self.n = True, self.n + 1

        def estimate_pose_finish(self):
            assert self.busy
            self.busy = False
            pose = np.eye(4, dtype=np.float32)
            pose[0, 3] = self.tag
            return pose, 7

    class LocalTickets:
        def __init__(self, n):
            self.it = iter(range(n))

        def next(self):
            return next(self.it, None)
    cloud = lambda v: synth.Cloud(np.full((1, 3), v, np.float32), np.zeros((1, 3), np.float32), np.zeros((1, 4), np.uint8))   # noqa: E731
    for n_pairs in (0, 1, 3, 4, 11):
        pairs = [(cloud(k), cloud(k)) for k in range(n_pairs)]
        for tickets in (None, LocalTickets(n_pairs)):
            ctxs = [Ctx() for _ in range(3)]
            res = sequence.alignPairs(ctxs, pairs, None, tickets=tickets)
            assert [r.pose[0, 3] for r in res] == [float(k) for k in range(n_pairs)]
            assert all(r.nIterations == 7 for r in res) and not any(c.busy for c in ctxs)
            assert sorted(c.n for c in ctxs) == sorted(len(range(j, n_pairs, 3)) for j in range(3))
