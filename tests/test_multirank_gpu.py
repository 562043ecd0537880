"""Engine-level data-parallel correctness (/root/reference/train.py:68,74 MirroredStrategy semantics): the
all-reduced gradient of a 2-rank train step equals the SUM of the two single-rank gradients (the 1/N average is folded into
clip + Adam), the two replicas end the step with identical weights, and those weights equal a single-process step that
applies the mean gradient.  NCCL when the box has two GPUs, gloo on CUDA tensors of one shared GPU otherwise (the bucketed /
asynchronous all-reduce path of engine.backward is the same)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, use_nccl, overlap, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import satk_path
    satk = satk_path.load()
    from importlib import import_module
    E = import_module("self-attention-tacotron_b200.engine")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SATK_AR_OVERLAP"] = "1" if overlap else "0"
    dev = torch.device("cuda", rank if use_nccl else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
    hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(7, "random")
    eng = E.TacotronEngine(hp, str(dev), params=ps)
    f, l = satk.synthetic_batch(hp, 6, 40, 48, seed=100 + rank, device=dev)
    masks = {k: v.to(dev) for k, v in satk.make_masks(d, 6, 40, 24, seed=200 + rank).items()}

    def allreduce(flat, async_op=False):
        return dist.all_reduce(flat, async_op=async_op)
    allreduce.supports_async = True
    # single-rank gradient of this rank's batch
    eng.forward(f, l, True, masks)
    eng.backward()
    torch.cuda.synchronize()
    own = eng.ps.grad.clone()
    # the data-parallel step
    eng.train_step(f, l, masks, allreduce=allreduce, world_size=world)
    torch.cuda.synchronize()
    torch.save(dict(own=own.cpu(), summed=eng.ps.grad.cpu(), flat=eng.ps.flat.cpu()), os.path.join(tmp, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_train_step_matches_single_rank_gradients(satk, root, tmp_path, overlap):
    import torch.multiprocessing as mp
    from importlib import import_module
    E = import_module("self-attention-tacotron_b200.engine")
    use_nccl = torch.cuda.device_count() >= 2
    port = 29500 + (os.getpid() % 2000) + (1 if overlap else 0)
    mp.spawn(_worker, args=(2, port, use_nccl, overlap, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(2))
    want = r0["own"] + r1["own"]
    scale = want.abs().max().item()
    for r in (r0, r1):     # all-reduce = sum of the per-replica gradients (same kernels, so up to the reduction order of the collective)
        assert (r["summed"] - want).abs().max().item() <= 1e-6 * scale
    assert torch.equal(r0["flat"], r1["flat"]), "replicas diverged"
    # a single process applying the mean gradient reaches the same weights
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    eng = E.TacotronEngine(hp, "cuda", params=satk.ParamStore(d).init(7, "random"))
    eng.ps.grad.copy_(want.cuda())
    eng.optimizer_step(world_size=2)
    torch.cuda.synchronize()
    assert (eng.ps.flat.cpu() - r0["flat"]).abs().max().item() <= 1e-7
