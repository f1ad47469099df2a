"""2-GPU (NCCL) test of the data-parallel head: DistributedSTiLHead on two ranks == the single-process oracle on
the concatenated batch (global-batch InfoNCE, all-reduced prototype partials).  Skipped with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, use_graph, transport):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import stil_tta_b200 as S
        from stil_tta_b200 import synth
        from oracle import stil_head_oracle as O
        cfg = synth.dvm_config(batch)
        batches = [synth.make_batch(cfg, seed=2022, rank=r) for r in range(world)]
        head = S.DistributedSTiLHead(cfg, device=f"cuda:{rank}", use_graph=use_graph, transport=transport)
        head.load(batches[rank])
        for _ in range(4):          # every run accumulates again; p2p alternates its two regions
            head.run()
        torch.cuda.synchronize()
        # oracle: reference CLIPLoss on the concatenation of all ranks' rows
        a = [b["feat_i"].float().requires_grad_(True) for b in batches]
        bb = [b["feat_t"].float().requires_grad_(True) for b in batches]
        loss, _, _ = O.clip_loss_global(a, bb, cfg.temperature, cfg.lambda_0)
        ga, gb = torch.autograd.grad(loss, (a[rank], bb[rank]))
        got = float(head.out["losses"][0])
        assert abs(got - float(loss)) <= 1e-3 * float(loss), (got, float(loss))
        for name, ref in (("d_feat_i", ga), ("d_feat_t", gb)):
            err = float((head.out[name].float().cpu() - ref).abs().max() / ref.abs().max())
            assert err <= 1e-3, (name, err)      # fp32 gradients of bf16 embeddings (the head's default)
        # row-local outputs == single-rank oracle; prototype partials == sum over ranks
        os_ = [O.head_step(b, cfg, with_grads=False) for b in batches]
        for k in ("max_idx", "mask1", "case1", "case3"):
            assert torch.equal(head.out[k].cpu(), os_[rank][k]), k
        assert abs(float(head.out["losses"][1]) - float(os_[rank]["loss_pt"])) <= 1e-3 * float(os_[rank]["loss_pt"]) + 1e-6
        cs = sum(o["class_sum"] for o in os_)
        cc = sum(o["class_count"] for o in os_)
        assert float((head.out["class_sum"].cpu() - cs).abs().max()) <= 1e-4
        assert float((head.prototypes_sum.cpu() - 4 * cs).abs().max()) <= 4e-4
        assert float((head.prototypes_count_sum.cpu() - 4 * cc).abs().max()) <= 4e-5
        ok = True
    except BaseException:
        import traceback
        traceback.print_exc()
        ok = False
    # a graph that captured NCCL work pins the communicator: drop it, then leave without the process-group teardown
    try:
        head.release()
        torch.cuda.synchronize()
    except BaseException:
        pass
    os._exit(0 if ok else 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("use_graph,transport", [(False, "fused"), (True, "fused"), (False, "p2p"), (True, "p2p"), (True, "nccl")])
def test_distributed_head_two_ranks(use_graph, transport):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), 256, use_graph, transport), nprocs=2, join=True)
