"""Two-GPU check (skipped on a single-GPU box): the view-sharded run returns, for every rank's views, bit-identical
pointmaps / mask logits to the single-GPU run, and identical replicated outputs (KAT C11) — for the
v1 head, the v2 head (InputMixer + LoftUp, whose batch-global MinMaxScaler needs the 3-channel min/max all-reduce,
model/upscalers/loftup.py:14-19) and a portrait scene, each with a ragged 3 + 2 view split."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, variant="v1", portrait=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import weights as W
        from oracle.panst3r import build_panst3r as build_oracle
        from helpers import bf16_weights
        from panst3r_b200.dist import ShardedPanSt3R, partition_views
        from panst3r_b200.panst3r import build_panst3r
        depth = (2, 2, 2) if variant == "v1" else (1, 1, 1, 1)
        sd = bf16_weights(W.synth_state_dict(build_oracle(variant, *depth), seed=3))
        m = build_panst3r(variant, *depth)
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        classes = [f"c{i}" for i in range(9)]
        m.panoptic_decoder.text_encoder.class_embeddings = W.synth_class_embeddings(classes)
        V, H, Wd = 5, 64, 96  # ragged split 3 + 2
        g = torch.Generator().manual_seed(1)
        imgs = (torch.rand(1, V, 3, H, Wd, generator=g) * 2 - 1).cuda()
        ts = torch.tensor([[[Wd, H] if portrait else [H, Wd]] * V])  # portrait: stored transposed (landscape tensors)
        pan_s, pm_s = ShardedPanSt3R(m, rank, world)(imgs, ts, classes)
        pan_1, pm_1 = m(imgs, ts, classes)
        s, e = partition_views(V, world)[rank]
        assert torch.equal(pm_s, pm_1[:, s:e]), "pointmaps differ"
        assert torch.equal(pan_s["pred_masks"], pan_1["pred_masks"][:, s:e]), "mask logits differ"
        assert torch.equal(pan_s["pred_logits"], pan_1["pred_logits"]) and torch.equal(pan_s["out_queries"], pan_1["out_queries"])
        for a, b in zip(pan_s["aux_outputs"], pan_1["aux_outputs"]):
            assert torch.equal(a["pred_masks"], b["pred_masks"][:, s:e])
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("variant,portrait", [("v1", False), ("v2", False), ("v1", True)])
def test_sharded_equals_single_gpu(variant, portrait):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + 7 * len(variant) + int(portrait)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, variant, portrait)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
