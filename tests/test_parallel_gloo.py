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


# ---------------------------------------------------------------------------------------------------------------------
# the training driver over two ranks: gradients packed into the flat bucket, all-reduced, and read by MultiTensorAdam
# straight from the bucket with the 1/world factor folded in (optimizer.exchange_and_step)
def _tiny_trainer(seed):
    import swapping_autoencoder_pytorch_b200 as S
    from swapping_autoencoder_pytorch_b200 import backend, default_options
    from tests.cpu_emulation import EmulatedKernels
    from oracle.fixtures import TINY
    backend.set_kernels(EmulatedKernels())
    torch.set_default_dtype(torch.float64)
    opt = default_options(**dict(TINY, R1_once_every=1))
    torch.manual_seed(seed)
    model = S.create_model(opt)
    return opt, model, S.create_optimizer(opt, model)


def _driver_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.fixtures import rnd
    opt, model, trainer = _tiny_trainer(100 + rank)          # different init per rank: rank 0's is broadcast
    assert trainer.world == 2 and model.defer_to_optimizer
    full = rnd(900, 4, 3, 64, 64).clamp(-1, 1)
    x = model.shard(full)
    torch.manual_seed(7)                                      # same crop / noise draws on both ranks and in the replay below
    trainer.train_one_step({"real_A": x}, 0)                  # D + R1
    trainer.train_one_step({"real_A": x}, 0)                  # G
    sd = {k: v.clone() for k, v in model.singlegpu_model.state_dict().items()}
    torch.save(sd, os.path.join(out, "driver_r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_training_driver_two_ranks_matches_manual_gradient_average(tmp_path):
    world = 2
    mp.spawn(_driver_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "driver_r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "driver_r1.pt"))
    for k in r0:
        assert torch.equal(r0[k], r1[k]), "ranks diverged on %s" % k
    # single-process replay: per-shard gradients averaged by hand, stock torch.optim.Adam
    from oracle.fixtures import rnd
    prev = torch.get_default_dtype()
    try:
        opt, model, trainer = _tiny_trainer(100)             # rank 0's initialisation
        inner = model.singlegpu_model
        full = rnd(900, 4, 3, 64, 64).clamp(-1, 1)
        shards = [full[:2], full[2:]]
        c = opt.R1_once_every / (1 + opt.R1_once_every)
        adam_d = torch.optim.Adam(trainer.Dparams, lr=opt.lr * c, betas=(opt.beta1 ** c, opt.beta2 ** c))
        adam_g = torch.optim.Adam(trainer.Gparams, lr=opt.lr, betas=(opt.beta1, opt.beta2))

        def averaged_step(params, frozen, adam, loss_fn, rng_states):
            trainer.set_requires_grad(frozen, False)
            trainer.set_requires_grad(params, True)
            grads, states_after = None, []
            for shard, state in zip(shards, rng_states):
                torch.set_rng_state(state)
                adam.zero_grad()
                loss_fn(shard).backward()
                states_after.append(torch.get_rng_state())
                g = [None if p.grad is None else p.grad.clone() for p in params]
                grads = g if grads is None else [a if b is None else a + b for a, b in zip(grads, g)]
            for p, g in zip(params, grads):
                p.grad = None if g is None else g / len(shards)
            adam.step()
            return states_after

        torch.manual_seed(7)
        s0 = [torch.get_rng_state()] * 2
        s1 = averaged_step(trainer.Dparams, trainer.Gparams, adam_d,
                           lambda x: sum(v.mean() for v in inner(x, command="compute_discriminator_losses")[0].values()), s0)
        s2 = averaged_step(trainer.Dparams, trainer.Gparams, adam_d,
                           lambda x: sum(v.mean() for v in inner(x, command="compute_R1_loss").values()) * opt.R1_once_every, s1)
        averaged_step(trainer.Gparams, trainer.Dparams, adam_g,
                      lambda x: sum(v.mean() for v in inner(x, None, None, command="compute_generator_losses")[0].values()), s2)
        ref = inner.state_dict()
        worst = max(float((r0[k].double() - ref[k].double()).abs().max()) for k in ref if ref[k].dtype.is_floating_point)
        assert worst < 1e-9, worst
    finally:
        torch.set_default_dtype(prev)
        from swapping_autoencoder_pytorch_b200 import backend
        backend.set_kernels(None)
