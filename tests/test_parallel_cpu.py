"""world_size-2 gloo test (CPU) of the bucketed gradient exchange: stage -> bucket mapping covers
the flat gradient buffer exactly once and the all-reduce averages across ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeModel:
    """Flat gradient buffer with the same bucket API as mmtg_b200.model.MMTG."""

    def __init__(self, nl, per_layer, tail, rank):
        self.nl, self.per, self.tail = nl, per_layer, tail
        n = nl * per_layer + tail
        self._flat = (None, None, torch.full((n,), float(rank + 1)))

    def layer_bucket(self, l):
        return l * self.per, (l + 1) * self.per

    def tail_bucket(self):
        return self.nl * self.per, self.nl * self.per + self.tail

    def tail_buckets(self):
        lo, hi = self.tail_bucket()
        mid = lo + self.tail // 3
        return (mid, hi), (lo, mid)


def _worker(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmtg_b200.parallel import GradSync
    m = _FakeModel(nl=4, per_layer=1000, tail=333, rank=rank)
    sync = GradSync()
    nstage = m.nl + 3
    touched = torch.zeros_like(m._flat[2])
    for s in range(nstage):
        before = m._flat[2].clone()
        sync.after_stage(m, s, nstage)
        touched += (m._flat[2] != before).float()
    sync.finish(m)
    expect = sum(r + 1 for r in range(world)) / world
    assert torch.allclose(m._flat[2], torch.full_like(m._flat[2], expect))
    assert (touched == 1).all(), "every element must be reduced exactly once"
    assert sync.bytes_reduced == m._flat[2].numel() * 4
    # ragged batches: SUM all-reduce of gradients pre-scaled by B_local / B_global
    from mmtg_b200.parallel import ragged_batch_scale
    b_local = 3 + 2 * rank  # 3 and 5 rows
    scale = ragged_batch_scale(b_local)
    assert abs(scale - b_local / 8.0) < 1e-9
    m2 = _FakeModel(nl=2, per_layer=10, tail=7, rank=rank)
    m2._flat[2].fill_(scale * (rank + 1))  # per-rank mean gradient (rank + 1), pre-scaled
    s2 = GradSync(average=False)
    for s in range(m2.nl + 3):
        s2.after_stage(m2, s, m2.nl + 3)
    want = (3 * 1 + 5 * 2) / 8.0  # gradient of the mean over the 8 concatenated rows
    assert torch.allclose(m2._flat[2], torch.full_like(m2._flat[2], want))
    # early all-reduce of the tied wte gradient (parallel.GradSync.after_stage): the head part is reduced after
    # stage 0, the type-embedding rows the embedding stage adds LATER are reduced with the tail bucket; the result
    # must equal the plain average of the final per-rank gradients
    E, V = 8, 200

    class _TiedModel(_FakeModel):
        data_config = {"max_seq_length": 220, "max_sent_length": 20}

        def wte_range(self):
            lo, hi = self.tail_bucket()
            return hi - V * E, hi, E

        def tail_buckets(self):
            lo, hi = self.tail_bucket()
            return (lo + 40, hi), (lo, lo + 40)  # (projector.. + wte), (encoder)

    m3 = _TiedModel(nl=3, per_layer=50, tail=40 + 24 + V * E, rank=rank)
    G3 = m3._flat[2]
    g = torch.Generator().manual_seed(100 + rank)
    final = torch.randn(G3.numel(), generator=g)            # this rank's final gradient
    wlo, whi, _ = m3.wte_range()
    type_part = torch.zeros_like(final)
    type_part[wlo:wlo + 12 * E] = torch.randn(12 * E, generator=g)  # rows 0..11 get type-embedding gradient
    G3.copy_(final - type_part)                               # state after stage 0: everything but the type rows
    s3 = GradSync()
    ns3 = m3.nl + 3
    for s in range(ns3):
        s3.before_stages(m3, s, s + 1, ns3)
        if s == m3.nl + 1:
            G3.add_(type_part)                                # the embedding stage adds into (cleared) wte rows
        s3.after_stage(m3, s, ns3)
    s3.finish(m3)
    both = [torch.empty_like(final) for _ in range(world)]
    dist.all_gather(both, final)
    assert torch.allclose(G3, sum(both) / world, atol=1e-6), (G3 - sum(both) / world).abs().max()
    dist.destroy_process_group()


def test_gradsync_gloo_world2():
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
