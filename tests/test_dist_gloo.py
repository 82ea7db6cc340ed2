"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: view partition, ragged row all-gather, and the
shard/replicate scheme of the head checked with the oracle modules standing in for the CUDA ones."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_views():
    from panst3r_b200.dist import partition_views
    assert partition_views(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert partition_views(5, 2) == [(0, 3), (3, 5)]
    assert partition_views(1, 2) == [(0, 1), (1, 1)]
    for V in range(1, 20):
        for w in (1, 2, 3, 4, 8):
            p = partition_views(V, w)
            assert p[0][0] == 0 and p[-1][1] == V and all(a[1] == b[0] for a, b in zip(p[:-1], p[1:]))
            assert max(e - s for s, e in p) - min(e - s for s, e in p) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from panst3r_b200.dist import gather_rows, partition_views
        torch.manual_seed(0)
        # (1) equal and ragged row gathers reproduce the full tensor in view order
        for V in (4, 5):
            full = torch.arange(V * 6, dtype=torch.float32).view(V, 3, 2)
            parts = partition_views(V, world)
            s, e = parts[rank]
            got = gather_rows(full[s:e].clone(), [b - a for a, b in parts])
            assert torch.equal(got, full), (V, rank)
        # (1b) LoftUp's batch-global MinMaxScaler across view shards: one MAX all-reduce of [-min, max]; a rank without
        #      views contributes the neutral element (model/upscalers/loftup.py:14-19)
        from panst3r_b200.dist import allreduce_minmax
        mine = torch.tensor([[-0.5 - rank, 0.25 + rank], [0.1 * rank, 0.9 - 0.1 * rank], [-1.0, 1.0]])
        got = allreduce_minmax(mine, "cpu")
        assert torch.equal(got, torch.tensor([[-1.5, 1.25], [0.0, 0.9], [-1.0, 1.0]])), got
        got = allreduce_minmax(mine if rank == 0 else None, "cpu")
        assert torch.equal(got, torch.tensor([[-0.5, 0.25], [0.0, 0.9], [-1.0, 1.0]])), got
        # (2) shard/replicate scheme of the head: per-view upscaler on the owning rank, gather of stride-16 features
        #     and mask features, replicated query decoder -> local slice equals the single-process result
        from helpers import CLASSES, build_oracle_head, head_inputs
        m = build_oracle_head("v1")
        V = 3
        feats, imgs, pos, ts = head_inputs(V, 32, 48, seed=5)
        with torch.no_grad():
            ref = m(feats, imgs, pos, ts, CLASSES)
            parts = partition_views(V, world)
            s, e = parts[rank]
            cat = torch.cat(feats, -1)[0, s:e]
            fpn, mask_f = m.upscaler((cat, imgs[0, s:e]), (32, 48))
            counts = [b - a for a, b in parts]
            fpn_all = gather_rows(fpn[0], counts)[None]
            mask_all = gather_rows(mask_f, counts)[None]
            cls = m.text_encoder(CLASSES)
            out = m.mask_transformer([fpn_all], mask_all, ts, cls)
        from helpers import relmax
        # batch-size dependent CPU GEMM blocking changes fp32 summation order: 1e-4 relative, not bitwise
        assert relmax(out["pred_masks"][:, s:e], ref["pred_masks"][:, s:e]) < 1e-4
        assert relmax(out["pred_logits"], ref["pred_logits"]) < 1e-4
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
