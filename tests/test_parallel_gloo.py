"""CPU, world_size 2 over gloo: the data-parallel wrapper averages the active parameter group's gradients at the
end of backward, leaves frozen parameters alone and reproduces the full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        from argparse import Namespace
        self.opt = Namespace(num_gpus=0)
        self.a = torch.nn.Linear(5, 4)
        self.b = torch.nn.Linear(4, 3)

    def per_gpu_initialize(self):
        pass

    def get_parameters_for_mode(self, mode):
        return list(self.a.parameters()) if mode == "generator" else list(self.b.parameters())

    def loss(self, x):
        return self.b(torch.tanh(self.a(x))).pow(2).mean(dim=1)

    def forward(self, *args, command=None):
        return getattr(self, command)(*args)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from swapping_autoencoder_pytorch_b200.parallel import MultiGPUModelWrapper
    torch.manual_seed(100 + rank)            # different init per rank: the wrapper must broadcast rank 0's
    model = _Toy().double()
    wrapped = MultiGPUModelWrapper(model.opt, model)
    full = torch.from_numpy(__import__("numpy").random.RandomState(0).standard_normal((8, 5)))
    x = wrapped.shard(full)
    # freeze group "a": only b's gradients may be exchanged
    for p in model.a.parameters():
        p.requires_grad_(False)
    wrapped(x, command="loss").mean().backward()
    res = {"b.w": model.b.weight.grad.clone(), "a.w.grad_is_none": model.a.weight.grad is None,
           "a.w": model.a.weight.detach().clone()}
    # second backward with the other group active
    for p in model.a.parameters():
        p.requires_grad_(True)
    for p in model.b.parameters():
        p.requires_grad_(False)
        p.grad = None
    wrapped(x, command="loss").mean().backward()
    res["a.w.grad"] = model.a.weight.grad.clone()
    res["b.w.grad_is_none"] = model.b.weight.grad is None
    if rank == 0:
        # single-process full-batch reference with the same (rank-0) parameters
        ref = _Toy().double()
        ref.load_state_dict(model.state_dict())
        ref.loss(full).mean().backward()
        res["ref.b.w"] = ref.b.weight.grad
        res["ref.a.w"] = ref.a.weight.grad
    torch.save(res, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_gradient_exchange_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0["a.w"], r1["a.w"]), "parameters were not broadcast from rank 0"
    assert r0["a.w.grad_is_none"] and r1["a.w.grad_is_none"] and r0["b.w.grad_is_none"]
    assert torch.allclose(r0["b.w"], r1["b.w"], atol=1e-14)
    assert torch.allclose(r0["b.w"], r0["ref.b.w"], atol=1e-12)
    assert torch.allclose(r0["a.w.grad"], r1["a.w.grad"], atol=1e-14)
    assert torch.allclose(r0["a.w.grad"], r0["ref.a.w"], atol=1e-12)
